"""
Serializers (parity: /root/reference/lib/utils/serializer.py:8-41).

``PytorchSerializer`` is the wire format of the TCP fallback path (``torch.save`` into a bytes buffer).  Loading uses
``weights_only=False`` explicitly: payloads contain schema dataclasses, and torch >= 2.6 would otherwise refuse them
(the reference needs TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1 for the same reason).  The fast in-box path never serialises.
"""
import io
import pickle

import torch


class PickleSerializer:
    @staticmethod
    def dumps(obj) -> bytes:
        return pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)

    @staticmethod
    def loads(buf: bytes):
        return pickle.loads(buf)


class JoblibSerializer:
    @staticmethod
    def dumps(obj) -> bytes:
        import joblib
        stream = io.BytesIO()
        joblib.dump(obj, stream)
        return stream.getvalue()

    @staticmethod
    def loads(buf: bytes):
        import joblib
        return joblib.load(io.BytesIO(buf))


class PytorchSerializer:
    @staticmethod
    def dumps(obj) -> bytes:
        stream = io.BytesIO()
        torch.save(obj, stream, pickle_protocol=pickle.HIGHEST_PROTOCOL)
        return stream.getvalue()

    @staticmethod
    def loads(buf: bytes):
        return torch.load(io.BytesIO(buf), weights_only=False)
