"""dgll_b200 — B200-native (sm_100a) neighbourhood-aggregation path for dke-lab/dgll.

``dgll_b200.kernels``   torch-tensor launchers over the C ABI (``include/dgll_b200.h``)
``dgll_b200.build``     in-tree nvcc build of ``libdgll_b200.so``
``dgll_b200._lib``      ctypes binding (no CPU fallback: a missing library raises)

The host-side mirror of the reference's operator/layer interface lives in
``dgll_b200.gcn_extension`` (dgll/FusedKernel/gcn_extension.cpp:103-110),
``dgll_b200.nn`` (dgll/nn) and ``dgll_b200.data`` (dgll/data, dgll/sampling,
dgll/dataloader, dgll/FeatureCache).
"""
__version__ = "0.1.0"
