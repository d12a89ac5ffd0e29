"""CPU model of the time-segment scheme behind tdm_process_long, built from the canonical-order
oracle only: overlapping segments demodulated independently from reset state, joined by content, unconverged segments
redone as the continuation of their predecessor.  It pins the SCHEME (what the CUDA orchestration in tdm_api.cu /
tdm_stitch.cu implements and tests/test_long_gpu.py checks on the device): the stitched dibits equal the sequential
chain's from its lock point on."""
import numpy as np
import pytest


def _run(O, iq):
    ob = O.OracleB(1)
    cb, _, db, _ = ob.process(iq[None])
    return ob, db[0, :cb[0]].copy()


def _find_join(prev, cur, K, lo, hi):
    tail = prev[-K:]
    hits = [j for j in range(max(lo, K), min(hi, len(cur)) + 1) if np.array_equal(cur[j - K:j], tail)]
    return hits[0] if len(hits) == 1 else -1


def _stitch(O, iq, S, W, K=None):
    N = len(iq)
    L = ((N - W) // S) & ~7
    K = K or min(max(W // 8, 128), 4096)
    runs = [_run(O, iq[s * L:(s + 1) * L + W]) for s in range(S)]           # (oracle with its final state, stream)
    out, redone = [runs[0][1]], 0
    prev_ob, prev_stream = runs[0]
    for s in range(1, S):
        ob, stream = runs[s]
        j = _find_join(prev_stream, stream, K, W // 2 - 2048, W // 2 + 256)
        if j >= 0:
            out.append(stream[j:])
            prev_ob, prev_stream = ob, stream
        else:                                                              # continue the predecessor over [sL + W, (s+1)L + W)
            cb, _, db, _ = prev_ob.process(iq[None, s * L + W:(s + 1) * L + W])
            cont = db[0, :cb[0]].copy()
            out.append(cont)
            prev_stream = cont                                             # prev_ob now sits at the end of this segment
            redone += 1
    covered = S * L + W
    if covered < N:
        cb, _, db, _ = prev_ob.process(iq[None, covered:])
        out.append(db[0, :cb[0]].copy())
    return np.concatenate(out), redone


@pytest.mark.parametrize("warmup,K,expect_redo", [(65536, None, 0), (32768, None, None), (10000, 4096, 4)])
def test_stitched_segments_equal_the_sequential_chain_after_lock(O, warmup, K, expect_redo):
    """65536: every join is found.  32768: whatever happens (some channels converge late), the result must hold.
    10000 with K = 4096: no segment can show 4096 agreeing dibits within 5000 symbols, so every one is redone as the
    continuation of its predecessor -- the boundary arithmetic of the redo path.  (With a short warm-up AND a short K
    the contract does not hold: a chain that has only just locked still makes a stray decision error now and then,
    which is why K is W/8 up to 4096 in the product.)"""
    N, S = 600_000, 5
    iq = O.generate(1, N, first_channel=1)[0]
    _, seq = _run(O, iq)
    got, redone = _stitch(O, iq, S, warmup, K)
    if expect_redo is not None:
        assert redone == expect_redo, redone
    assert abs(len(got) - len(seq)) <= 1, (len(got), len(seq))
    n = min(len(got), len(seq))
    tx = O.tx_dibits(1, N // 2)
    lag = min(range(10, 30), key=lambda d: int((seq[d + 200_000:d + 210_000] != tx[200_000:210_000]).sum()))
    bad = np.flatnonzero(seq[lag:] != tx[:len(seq) - lag])
    lock = (int(bad.max()) + 1 + lag) if len(bad) else 0
    lock = max(lock, (N - warmup) // S // 2)          # later segments start from states that had a whole segment to settle
    assert lock < 150_000
    assert np.array_equal(got[lock:n], seq[lock:n])
