"""yue_b200 -- B200-native BPR hot path of 0411tony/Yue (triplet SGD epoch + full-catalog
masked top-N ranking; APR and WRMF on the same tables) behind Yue's recommender class API.  See DESIGN.md.

Importing the package is cheap; the CUDA library is loaded on first use of Engine and there is
no CPU fallback (yue_b200._lib.load raises if libyue_b200.so is missing)."""
__version__ = "0.1.0"
