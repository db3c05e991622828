"""telluride_decoding_b200: the dense decoding hot path of google/telluride_decoding on B200.

Lagged (Hankel) Gram / cross-covariance accumulation, ridge and CCA solves, windowed correlation
and the attended-speaker decision, behind the reference's brain_data / brain_model / cca /
regression / infer_decoder Python API.  Python + ctypes over hand-written sm_100a CUDA
(libtdgram.so, include/tdgram.h); torch is used for device memory and DLPack only.
"""

__version__ = '0.1.0'
