#!/bin/bash
# re-entry verification: full GPU suite (timed), HuBERT + infer_panel tools
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1; tail -6 gpurun_out/c1_pytest.log
timeout 200 python tools/bench_hubert.py > gpurun_out/bench_hubert.json 2> gpurun_out/bench_hubert.err; cat gpurun_out/bench_hubert.json; tail -3 gpurun_out/bench_hubert.err
for gr in 1 0; do
  EVK_INFER_GRAPH=$gr timeout 200 python tools/bench_infer_panel.py > gpurun_out/bench_infer_panel_g$gr.json 2> gpurun_out/bench_infer_panel.err; cat gpurun_out/bench_infer_panel_g$gr.json; tail -3 gpurun_out/bench_infer_panel.err
done
