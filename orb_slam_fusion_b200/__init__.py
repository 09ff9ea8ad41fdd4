"""orb_slam_fusion_b200 -- the ORB front end of J094/orb_slam_fusion (OrbExtractor::operator(), the
Hamming matching kernels and the bag-of-words transform) as hand-written sm_100a CUDA behind a C ABI (include/orbx.h).

Only what the hot path needs lives here: csrc/ (kernels + C ABI) and the host-side mirror of the
reference's OrbExtractor / ORBmatcher interface.  There is no CPU fallback."""
from ._abi import KP_DTYPE, WQ_DTYPE, WR_DTYPE, OrbxError, LIB_PATH, OPT_CLAIM_SEQUENTIAL  # noqa: F401
from .orb_extractor import OrbExtractor, synth_frames  # noqa: F401
from .orb_matcher import ORBmatcher, synth_descriptors, popc_peak  # noqa: F401
from .orb_vocabulary import ORBVocabulary  # noqa: F401
