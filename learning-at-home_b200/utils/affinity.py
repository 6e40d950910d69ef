"""CPU / NUMA placement helper for one-process-per-GPU launches (no counterpart in the reference, which targets volunteers'
PCs): pinned staging buffers are placed by first touch, so a rank should run next to its GPU before it allocates them."""
import os


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Restrict this process to the CPUs NVML reports as local to GPU ``local_rank`` (honours CUDA_VISIBLE_DEVICES).
    Best effort: returns False when NVML is unavailable or the container forbids affinity changes."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = local_rank
        if visible:
            ids = visible.split(",")
            if local_rank < len(ids) and ids[local_rank].strip().isdigit():
                index = int(ids[local_rank])
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return True
    except Exception:
        return False
