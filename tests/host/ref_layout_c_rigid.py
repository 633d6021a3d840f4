"""`Rigid/c_rigid` of the reference's install layout (CMakeLists.txt:24-27 installs the compiled
module there), served by this repository's pybind11 host class.  Copied by
oracle/stage_ref_tests.sh next to the reference's own, unmodified __init__.py and Rigid.py so that
`from Rigid import c_rigid` (src/Rigid.py:1) binds the B200 implementation."""
from rigid_body_light_b200.c_rigid import CManyBodies, host_class, precision  # noqa: F401
