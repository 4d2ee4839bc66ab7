"""Response object printed on the stdout protocol (mirror of /root/reference/src/utils/response/__init__.py:17-33)."""


class ResponseStatus:
    SUCCESS = "success"
    FAILED = "failed"


class EaseVoiceResponse:
    def __init__(self, status, message, data=None, uuid=None):
        self.status, self.message, self.data, self.uuid = status, message, data, uuid

    def to_dict(self):
        return {"status": self.status, "message": self.message, "data": self.data, "uuid": self.uuid}
