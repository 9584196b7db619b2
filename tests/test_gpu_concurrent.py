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
