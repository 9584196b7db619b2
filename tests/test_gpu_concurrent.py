"""GPU: several contexts in flight on one device (what bench.py --in-flight and the CLI's --in-flight do):
every context, driven by its own host thread at the same time, must return exactly what a context
running alone returns -- no shared scratch, graph or stream state between wb_ctx objects."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PROMPT = [50258, 50259, 50359, 50363]
EOT = 50257


def test_contexts_in_flight_match_a_solo_run(wb):
    B, S, n_new = 16, 3, 24
    pcm = wb.synth.batch(B, seed=41, seconds=30.0)
    solo = wb.Whisper(wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=B, max_chunks=B))
    want = solo.transcribe_batch(list(pcm), PROMPT, n_new, EOT)[0]
    solo.close()
    ctxs = [wb.Whisper(wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=B, max_chunks=B)) for _ in range(S)]
    got = [[] for _ in range(S)]
    errs = []

    def work(i):
        try:
            for _ in range(4):
                got[i].append(ctxs[i].transcribe_batch(list(pcm), PROMPT, n_new, EOT)[0])
        except Exception as e:          # surfaced below, a thread must not die silently
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in ctxs:
        c.close()
    assert not errs, errs
    for i in range(S):
        assert len(got[i]) == 4
        for g in got[i]:
            assert g == want


def test_contexts_created_while_others_capture_and_decode(wb):
    """What the CLI's --in-flight workers do: every thread creates ITS OWN context and starts right away, so one
    thread's wb_create (allocation, uploads) overlaps another's first decode (CUDA-graph capture).  Nothing on
    that path may be a device-wide operation (cudaDeviceSynchronize is illegal while any stream captures)."""
    B, S, n_new = 8, 4, 6
    pcm = wb.synth.batch(B, seed=43, seconds=30.0)
    out, errs = [None] * S, []

    def work(i):
        try:
            for rep in range(2):               # second round: fresh contexts while the others are mid-flight
                m = wb.Whisper(wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=B, max_chunks=B))
                for n in (n_new, n_new + 1 + i):          # a second shape = a second capture later on
                    r = m.transcribe_batch(list(pcm), PROMPT, n, EOT)[0]
                    if n == n_new:
                        out[i] = r
                m.close()
        except Exception as e:
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert all(o == out[0] for o in out)
