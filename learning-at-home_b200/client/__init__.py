from .remote_expert import RemoteExpert, RemoteExpertError
from .gating_function import GatingFunction

__all__ = ["RemoteExpert", "GatingFunction", "RemoteExpertError"]
