"""Writer side of the parent/child stdout line protocol (reference: src/utils/helper/connector.py:35-71).
The REST service launches the trainer as a subprocess and parses these prefixed JSON lines; the prefixes and
payload keys are the wire format and are kept byte-identical."""
import json

RESP_PREFIX = "response-of-easevoice"
LOSS_PREFIX = "loss-of-easevoice"
LOG_PREFIX = "log-of-easevoice"
SESSION_PREFIX = "session-data-of-easevoice"


class MultiProcessOutputConnector:
    def _print(self, prefix, data):
        print(f"{prefix} {data}", flush=True)

    def write_response(self, resp):
        self._print(RESP_PREFIX, json.dumps(resp.to_dict()))

    def write_session_data(self, data):
        self._print(SESSION_PREFIX, json.dumps(data))

    def write_loss(self, step, loss, other=None):
        data = {"step": step, "loss": loss}
        if other is not None:
            data.update(other)
        self._print(LOSS_PREFIX, json.dumps(data))

    def write_log(self, log):
        self._print(LOG_PREFIX, json.dumps(log))
