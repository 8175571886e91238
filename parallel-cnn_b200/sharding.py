"""Host-side statement of the sample sharding the fused kernels apply on the device (csrc/fused_kernels.cu, k_fused
prologue and effective_global_batch): rank r of `world` takes B consecutive samples starting at cursor + r * B, the
global batch is clamped at the end of the split, and the step size is dt / (effective global batch).  No collective
is involved in the data path; the only exchange per step is the all-reduce of the 2,344-float packed gradient."""


def shard(cursor, B, rank, world, n):
    """(first sample, count) this rank processes in the step that starts at global position `cursor`."""
    base = cursor + rank * B
    nb = max(0, min(B, n - base))
    return base, nb


def effective_global_batch(cursor, B, world, n):
    return max(1, min(B * world, n - cursor))


def steps_per_epoch(n, B, world):
    gb = B * world
    return (n + gb - 1) // gb


def next_cursor(cursor, B, world, n):
    c = cursor + B * world
    return 0 if c >= n else c
