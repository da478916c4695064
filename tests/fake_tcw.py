"""A stand-in for ``pyfstat.tcw_fstat_map_funcs`` with the same registry / dispatcher
interface (names, argument meaning, error behaviour: tcw:320-327, 351-358, 361-492, 495-544),
for boxes where neither PyFstat nor /root/reference exists (the GPU box).  Written for the
tests; the CPU suite runs the same assertions against the real reference module.
"""

from time import time

fstatmap_versions = {
    "lal": lambda multiFstatAtoms, windowRange, BtSG: (_ for _ in ()).throw(RuntimeError("no lal here")),
}


def _get_transient_fstat_map_features():
    return {"lal": False, "pycuda": False}


def init_transient_fstat_map_features(feature="lal", cudaDeviceName=None):
    features = _get_transient_fstat_map_features()
    if feature == "pycuda":
        raise RuntimeError("pycuda use was requested, but imports failed.")
    elif feature == "lal":
        gpu_context = None
    else:
        raise ValueError(f"Unknown transient F-stat map computation feature'{feature}' requested.")
    return features, gpu_context


def call_compute_transient_fstat_map(version, features, multiFstatAtoms=None, windowRange=None, BtSG=False):
    if version in fstatmap_versions:
        if features[version]:
            time0 = time()
            FstatMap = fstatmap_versions[version](multiFstatAtoms, windowRange, BtSG)
            timingFstatMap = time() - time0
        else:
            raise Exception('Required module(s) for transient F-stat map method "{}" not available!'.format(version))
    else:
        raise Exception('Transient F-stat map method "{}" not implemented!'.format(version))
    return FstatMap, timingFstatMap
