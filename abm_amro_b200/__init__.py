"""amro_b200 -- B200-native implementation of ABM_AMRO's ward-level colonisation path.

* ``abm_amro_b200.amro``      the drop-in module (same seven functions as the reference's ``amro``)
* ``abm_amro_b200.engine``    handle-based API over the C ABI (compile a hospital once, simulate many times)
* ``abm_amro_b200.capi``      ctypes binding of include/amro_b200.h
* ``abm_amro_b200.synth``     synthetic hospitals of the BASELINE.json shapes
* ``abm_amro_b200.ensemble``  sharding of an ensemble over the GPUs of one box (torch.distributed plumbing)

The compiled pieces are built in-tree by ``__graft_entry__.build()``; importing ``amro`` without them
raises (there is no CPU fallback).
"""
__version__ = "0.2.2"


def install_as_amro():
    """Make ``import amro`` resolve to this implementation (drop-in for scripts written for the reference)."""
    import sys
    from . import amro as _amro
    sys.modules["amro"] = _amro
    return _amro
