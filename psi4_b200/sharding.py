"""Q-sharding rule shared by the host mirror and libb200jk.so (engine.cu, b200jk_set_layout):
shard r of w owns auxiliary rows [naux*r//w, naux*(r+1)//w) -- contiguous, near-equal.
J and K are sums over Q (the reference's own Q-block loop, dfhelper.cc:3124-3159, is this
decomposition executed serially with beta=1), so partial results combine with one sum."""


def q_range(naux: int, rank: int, world: int) -> tuple[int, int]:
    return (naux * rank) // world, (naux * (rank + 1)) // world
