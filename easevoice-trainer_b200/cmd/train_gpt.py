#!/usr/bin/env python
"""Subprocess entry point with the reference's CLI contract (src/cmd/train_gpt.py:20-43):
    python train_gpt.py -c <json of GPTTrainParams>
prints exactly one `response-of-easevoice {...}` line at the end; periodic `loss-of-easevoice {...}` lines."""
import argparse
import json
import os
import sys
import traceback
from dataclasses import asdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    from easevoice_trainer_b200.train.gpt import GPTTrain, GPTTrainParams
    from easevoice_trainer_b200.utils.connector import MultiProcessOutputConnector
    from easevoice_trainer_b200.utils.response import EaseVoiceResponse, ResponseStatus
    connector = MultiProcessOutputConnector()
    try:
        ap = argparse.ArgumentParser(description="run train gpt")
        ap.add_argument("-c", "--config", type=argparse.FileType("r"), required=True)
        args = ap.parse_args()
        params = GPTTrainParams(**json.loads(args.config.read()))
        out = GPTTrain(params=params).train()
        connector.write_response(EaseVoiceResponse(ResponseStatus.SUCCESS, "Finish train gpt", data=asdict(out)))
    except Exception as e:
        traceback.print_exc()
        connector.write_response(EaseVoiceResponse(ResponseStatus.FAILED, f"failed to train gpt, {e}"))


if __name__ == "__main__":
    main()
