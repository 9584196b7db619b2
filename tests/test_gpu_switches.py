"""GPU: kernel variants selected by process-wide environment switches (read once per process, so each runs in a fresh
interpreter): the first-generation encoder attention, the FMA-pipe exponentials, the exponentials-before-P-wait ordering,
other grids of the vocabulary kernel, early / attention-wide release of programmatic-launch dependents.  Each variant runs
__graft_entry__.smoke(): log-mel <= 1e-4 of the C oracle, fp32 toy build token-identical to the oracle, bf16 whisper-base
encoder within 2e-2 and teacher-forced logits within 5e-2 of the fp32 oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"WB_ATTN_V": "1"},
    {"WB_ATTN_POLY": "2"},
    {"WB_ATTN_LATE": "1"},
    {"WB_VOCAB_CTAS": "37", "WB_PDL_LATE": "0"},
    {"WB_PDL_LATE": "2", "WB_VOCAB_L2": "none"},
]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: "+".join(f"{k}={v}" for k, v in sorted(e.items())))
def test_switch_variant_passes_smoke(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "smoke ok" in r.stdout
