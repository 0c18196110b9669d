"""poi_b200 -- B200-native training engine behind the reference's prog_*.py / public/*.py surface.

Layout (only what the hot path needs):
  csrc/      hand-written sm_100a CUDA kernels + the C-ABI (include/poi_engine.h)
  _lib.py    ctypes binding of libpoi_b200.so (fails loudly when it is not built)
  engine.py  thin Python wrapper around the C-ABI (raw device pointers of torch tensors)
  shared.py  theano.shared look-alike (.get_value / .set_value / .eval) over device tensors
  public/    the reference's model-class surface (GRU, GRU_Spatial, BPR, PRME, GeoIE, Valuate, ...)
  prog_*.py  the reference's drivers, ported to Python 3
"""
__all__ = ["public"]
