"""Frame / carrier sharding across the GPUs of one box.

Every frame (and every carrier channel) of the path is independent: the NCO restarts per
frame (carrier.rs:10-15), the FIR history starts at zero (fir.rs:13) and the Philox noise
stream is indexed by the GLOBAL frame id.  So the work is cut into contiguous frame ranges,
no sample ever crosses GPUs, and the only exchange is one all-reduce of the error counters.
"""


def shard_range(total, rank, world):
    """Contiguous range [start, start+count) of `total` items owned by `rank` of `world`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    start = total * rank // world
    return start, total * (rank + 1) // world - start


def shard_channels(n_channels, frames_per_channel, rank, world):
    """Whole channels per rank: returns (first channel, channel count, first global frame, frame count)."""
    c0, nc = shard_range(n_channels, rank, world)
    return c0, nc, c0 * frames_per_channel, nc * frames_per_channel
