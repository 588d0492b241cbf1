"""CPU test of the tile-DAG Cholesky kernel's job list (LumpCholKernel.cu buildJobs, through the C ABI hook
bspb200_lumpchol_job_list; no device needed). The kernel hands the jobs out by an arrival ticket in list order and a CTA
that holds a job spins until its inputs are published, so freedom from deadlock under ANY number of resident CTAs needs:
every job depends only on jobs EARLIER in the list, or on the chain CTA, which at step d in turn only waits for the two
hand-over jobs of block d and (transitively) for earlier chain steps. The test replays that argument: it walks the list
with a simulated chain and checks every dependency (reference algorithm replaced: cusolverDnDpotrf + cublasDtrsm on a
lump column, MatOpsCuda.cu:508-566 - a library call there, no schedule to check)."""
import ctypes as C

import numpy as np
import pytest

import baspacho_b200 as bsp


def job_list(nbc, nbr, seg, lag):
    api = bsp.api()
    n = api.lumpchol_job_list(nbc, nbr, seg, lag, None, 0)
    assert n > 0
    out = np.zeros((n, 5), dtype=np.int32)
    assert api.lumpchol_job_list(nbc, nbr, seg, lag, out.ctypes.data_as(C.c_void_p), n) == n
    return out


@pytest.mark.parametrize("nbc,nbr", [(4, 4), (5, 9), (11, 11), (21, 21), (7, 60), (55, 55)])
@pytest.mark.parametrize("seg,lag", [(0, 0), (1, 0), (3, 0), (4, 2), (8, 6)])
def test_job_list_is_complete_and_every_dependency_precedes(nbc, nbr, seg, lag):
    jobs = job_list(nbc, nbr, seg, lag)
    pos_final = {}      # (i, c, type) -> list position of the job that finishes the tile
    segs = {}           # (i, c, type) -> [(k0, k1, position)]
    for p, (i, c, k0, k1, tl) in enumerate(jobs):
        typ, last = tl & 3, tl >> 4
        segs.setdefault((i, c, typ), []).append((k0, k1, p))
        if last:
            assert (i, c, typ) not in pos_final
            pos_final[(i, c, typ)] = p
    # completeness: every tile of the trapezoid exactly once, its K blocks [0, K) covered by consecutive segments in order
    expect = set()
    for d in range(nbc):
        expect.add((d, d - 1, 2))
        if d > 0:
            expect.add((d, d - 1, 1))
    for c in range(nbc):
        for i in range(c + 2 if c + 1 < nbc else c + 1, nbr):
            expect.add((i, c, 0))
    assert set(segs) == expect == set(pos_final)
    for (i, c, typ), lst in segs.items():
        K = max(0, c)
        assert [s[0] for s in lst] == [0] + [s[1] for s in lst[:-1]] and lst[-1][1] == K
        assert [s[2] for s in lst] == sorted(s[2] for s in lst)          # a tile's segments in list order
        assert lst[-1][2] == pos_final[(i, c, typ)]
        assert len(lst) <= 256

    # who publishes L(r, k): the chain CTA for k = r - 1 (r < nbc), else the finishing job of the regular tile (r, k)
    def source(r, k):
        return ("chain", r) if (k == r - 1 and r < nbc) else ("job", pos_final[(r, k, 0)])

    # chain step d needs the hand-overs of block d; a job may need chain steps. Earliest list position after which chain
    # step d can complete = max over the hand-over jobs of blocks <= d (the chain is sequential)
    chain_ready = []
    worst = -1
    for d in range(nbc):
        worst = max(worst, pos_final[(d, d - 1, 2)], pos_final[(d, d - 1, 1)] if d > 0 else -1)
        chain_ready.append(worst)
    for p, (i, c, k0, k1, tl) in enumerate(jobs):
        typ, last = tl & 3, tl >> 4
        brow = i if typ == 2 else max(c, 0)
        deps_chain = -1
        for k in range(k0, k1):
            for r in (i, brow):
                kind, v = source(r, k)
                if kind == "job":
                    assert v < p, ("reads a tile finished later in the list", (i, c, typ, k0, k1), (r, k))
                else:
                    deps_chain = max(deps_chain, v)
        if last and typ == 0:
            deps_chain = max(deps_chain, c)   # the triangular product needs W_c = chain step c
        if deps_chain >= 0:
            # the chain step this job waits for only needs hand-overs that come EARLIER in the list than this job
            assert chain_ready[deps_chain] < p, ("waits for a chain step that waits for a later job", (i, c, typ), deps_chain)
    # hand-over jobs of block d never wait for chain step d or later
    for d in range(nbc):
        for typ in ((1, 2) if d > 0 else (2,)):
            for k0, k1, p in segs[(d, d - 1, typ)]:
                assert k1 <= max(0, d - 1)
