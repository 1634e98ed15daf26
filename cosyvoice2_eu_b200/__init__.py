"""cosyvoice2_eu_b200 -- B200-native (sm_100a) token2wav engine for CosyVoice2-EU.

Only the hot path lives here: csrc/ (hand-written CUDA kernels + the C ABI of include/cv2eu_b200.h) and the Python
host that mirrors the reference's interface for the path (flow.inference / hift.inference / token2wav)."""
from .engine import (B200Encoder, B200Estimator, B200Flow, B200HiFT, B200Token2Wav, GraphedToken2Wav, StreamGroup, euler_schedule,  # noqa: F401
                     get_engine)
from .frontend import (align_prompt, extract_speech_feat, extract_speech_feat_batch, extract_spk_feat,  # noqa: F401
                       mel_spectrogram, resample_16k_to_24k)
from .scheduler import StreamScheduler  # noqa: F401
from .lib import LIB_PATH, SYMBOLS, Cv2Error  # noqa: F401
