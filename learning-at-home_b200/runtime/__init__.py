from .task_pool import TaskPool, TaskPoolBase, Task
from .expert_backend import ExpertBackend
from .runtime import TesseractRuntime

__all__ = ["TaskPool", "TaskPoolBase", "Task", "ExpertBackend", "TesseractRuntime"]
