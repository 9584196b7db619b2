"""whisper-rust-ort_b200 — B200-native hot path (log-mel -> encoder -> greedy decode) behind the
C ABI of include/whisper_b200.h.  Python here is host glue only: ctypes binding, weight spec,
synthetic clips.  Import via the repo-root shim:  `import wb200`."""
from . import binding, synth, weights  # noqa: F401
from .binding import Whisper, Pool, WbError, default_cfg, lib, load_audio, WB_PREC_BF16, WB_PREC_FP32  # noqa: F401
