"""rigid_body_light_b200 -- B200-native (sm_100a) hot path of Rigid_Body_Light.

Layers:  Rigid.RigidBody (API mirror of /root/reference/src/Rigid.py)
-> c_rigid.CManyBodies (pybind11 C++ host class) -> C-ABI (include/rbl.h,
librbl.so) -> hand-written CUDA kernels.  There is no CPU fallback: importing the
host class without the built CUDA library raises.
"""
__version__ = "0.1.0"
