"""One-off GPU measurements behind the numbers in DESIGN.md / profiles/ (not the bench contract; see bench.py).

  python tools/gpu_probe.py decode [B] [new_tokens]   cluster-chained decoder layers vs the per-kernel path:
                                                      token agreement, ms per decode, ms per step
  python tools/gpu_probe.py inflight [S...]           throughput with S batches of 32 in flight
  python tools/gpu_probe.py kernels                   cross_attn / vocab_proj / logmel replays (wb_bench_kernel)
  python tools/gpu_probe.py mel [n_clips]             BASELINE.json configs[1]: K1a and K1a+K1b per kernel variant
"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import wb200  # noqa: E402

PROMPT, EOT = [50258, 50259, 50359, 50363], 50257


def make(B, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        m = wb200.Whisper(wb200.default_cfg("base", precision=wb200.WB_PREC_BF16, max_batch=B, max_chunks=B))
        if os.environ.get("WB_PROBE_HINT"):                  # wb_set_load_hint: 1 = latency-oriented kernels
            m.set_load_hint(int(os.environ["WB_PROBE_HINT"]))
        return m
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def decode(argv):
    B = int(argv[0]) if argv else 32
    new = int(argv[1]) if len(argv) > 1 else 128
    pcm = wb200.synth.fast_batch(B, seed=1)
    out = {}
    toks = {}
    for name, env in (("cluster", {"WB_DEC_CLUSTER": "1"}), ("per_kernel", {"WB_DEC_CLUSTER": "0"})):
        m = make(B, **env)
        m.upload_pcm(pcm)
        m.run_log_mel()
        m.encode(None, 0, B, want_hidden=False)
        best = 1e9
        for _ in range(4):
            t = m.greedy_decode(B, PROMPT, new, EOT)
            tm = m.timing()
            best = min(best, tm["decode_ms"])
        toks[name] = t
        out[name] = {"decode_ms": best, "ms_per_step": best / (len(PROMPT) + new - 1), "launches": tm["decode_launches"],
                     "steps": tm["decode_steps"]}
        m.close()
    a, b = np.array(toks["cluster"]), np.array(toks["per_kernel"])
    same = (a == b)
    first_diff = [int(np.argmax(~r)) if not r.all() else -1 for r in same]
    out["token_agreement"] = {"frac": float(same.mean()), "rows_identical": int(same.all(1).sum()), "rows": B,
                              "first_diff_pos_per_row": first_diff}
    out["head"] = {"cluster": toks["cluster"][0][:12], "per_kernel": toks["per_kernel"][0][:12]}
    print(json.dumps(out, indent=1))


def inflight(argv):
    counts = [int(x) for x in argv] or [1, 2, 4]
    B = int(os.environ.get("WB_PROBE_B", "32"))
    pcm = wb200.synth.fast_batch(B, seed=1)
    out = {"B": B}
    for S in counts:
        ctxs = [make(B) for _ in range(S)]
        for c in ctxs:
            c.upload_pcm(pcm)
        def work(i, n):
            for _ in range(n):
                ctxs[i].transcribe_resident(B, PROMPT, 128, EOT)
        for phase, n in (("warm", 2), ("timed", 6)):
            th = [threading.Thread(target=work, args=(i, n)) for i in range(S)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
        out[S] = {"audio_s_per_s": S * 6 * B * 30 / dt, "ms_per_batch": 1000 * dt / (S * 6), "stage_ms": ctxs[0].timing()}
        for c in ctxs:
            c.close()
    print(json.dumps(out, indent=1))


def stages(argv):
    """S contexts in flight, each looping ONE stage (decode only / encoder only): which stage saturates the GPU, and where?"""
    counts = [int(x) for x in argv] or [1, 8]
    B = int(os.environ.get("WB_PROBE_B", "32"))           # per-chain batch: how much of a decode is per-clip, how much per-step?
    pcm = wb200.synth.fast_batch(B, seed=1)
    out = {"B": B}
    ctxs = [make(B) for _ in range(max(counts))]
    for c in ctxs:
        c.upload_pcm(pcm)
        c.run_log_mel()
        c.encode(None, 0, B, want_hidden=False)
        c.greedy_decode(B, PROMPT, 128, EOT)
    for stage in ("decode", "encode"):
        for S in counts:
            n = 6 if stage == "decode" else 30
            def work(i):
                for _ in range(n):
                    if stage == "decode":
                        ctxs[i].greedy_decode(B, PROMPT, 128, EOT)
                    else:
                        ctxs[i].encode(None, 0, B, want_hidden=False)
                ctxs[i].timing()          # waits for the last enqueued stage
            th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
            out[f"{stage} S={S}"] = {"ms_per_batch": 1000 * dt / (S * n)}
    print(json.dumps(out, indent=1))
    for c in ctxs:
        c.close()


def kernels(argv):
    B = 32
    m = make(B)
    m.upload_pcm(wb200.synth.fast_batch(B, seed=1))
    m.run_log_mel()
    m.encode(None, 0, B, want_hidden=False)
    m.greedy_decode(B, PROMPT, int(argv[0]) if argv else 64, EOT)
    out = {}
    for k, it in (("cross_attn", 30), ("vocab_proj", 20), ("vocab_tc", 20), ("dec_vocab", 20), ("logmel", 10)):
        try:
            ms, by = m.bench_kernel(k, B, it)
        except wb200.WbError as e:          # a path that is switched off in this process
            out[k] = {"ms": float("nan"), "bytes": 0.0, "GBps": float("nan"), "skipped": str(e)}
            continue
        out[k] = {"ms": ms, "bytes": by, "GBps": by / ms / 1e6}
    for b in ([int(x) for x in os.environ["WB_PROBE_B"].split(",")] if "WB_PROBE_B" in os.environ else (1, 4, 7, 14, 28, 32)):
        try:
            ms, by = m.bench_kernel("dec_layers", b, 20)
        except wb200.WbError:
            break
        out[f"dec_layers_B{b}"] = {"ms": ms, "bytes": by, "GBps": by / ms / 1e6}
    print(json.dumps(out, indent=1))
    m.close()


def mel(argv):
    n = int(argv[0]) if argv else 1024
    pcm = wb200.synth.fast_batch(n, seed=1)
    out = {}
    for nm in (80, 128):
        cfg = wb200.default_cfg("toy", precision=wb200.WB_PREC_BF16, max_batch=4, max_chunks=n)
        cfg.n_mels = nm
        m = wb200.Whisper(cfg)
        m.upload_pcm(pcm)
        for packed in ("1", "0"):
            for tpc in ("3", "2"):
                os.environ["WB_MEL_PACKED"], os.environ["WB_MEL_CTAS_PER_SM"] = packed, tpc
                best = 1e9
                for _ in range(4):
                    m.run_log_mel()
                    best = min(best, m.timing()["mel_ms"])
                k1a, by = m.bench_kernel("logmel", 1, 5)
                out[f"n_mels={nm} packed={packed} ctas_per_sm={tpc}"] = {
                    "K1a+K1b_ms": best, "K1a_ms": k1a, "K1a_GBps": by / k1a / 1e6,
                    "pipeline_GBps": n * (1.92e6 + nm * 3000 * 4) / best / 1e6, "audio_s_per_s": n * 30 / (best * 1e-3)}
        os.environ.pop("WB_MEL_PACKED"); os.environ.pop("WB_MEL_CTAS_PER_SM")
        m.close()
    print(json.dumps(out, indent=1))


def xconc(argv):
    """cross_attn / vocab_proj replayed from S contexts at once: does the aggregate stream faster than one kernel alone?"""
    counts = [int(x) for x in argv] or [1, 2, 4, 8]
    B = 32
    pcm = wb200.synth.fast_batch(B, seed=1)
    out = {}
    ctxs = [make(B) for _ in range(max(counts))]
    for c in ctxs:
        c.upload_pcm(pcm)
        c.run_log_mel()
        c.encode(None, 0, B, want_hidden=False)
        c.greedy_decode(B, PROMPT, 8, EOT)
    for pdl in ("0", "1"):
        if pdl == "1":
            os.environ["WB_BENCH_PDL"] = "1"
        else:
            os.environ.pop("WB_BENCH_PDL", None)
        for kern in ("cross_attn", "vocab_proj"):
            for S in counts:
                res = [None] * S
                def work(i):
                    res[i] = ctxs[i].bench_kernel(kern, B, 300)
                th = [threading.Thread(target=work, args=(i,)) for i in range(S)]
                t0 = time.perf_counter()
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                dt = time.perf_counter() - t0
                tot = sum(r[1] for r in res) * 300
                out[f"{kern} pdl={pdl} S={S}"] = {"aggregate_GBps_wall": tot / dt / 1e9, "ms_per_launch_per_ctx": [round(r[0], 5) for r in res],
                                                 "aggregate_GBps_events": sum(r[1] / r[0] / 1e6 for r in res)}
    print(json.dumps(out, indent=1))
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    {"decode": decode, "inflight": inflight, "kernels": kernels, "mel": mel, "xconc": xconc, "stages": stages}[sys.argv[1]](sys.argv[2:])
