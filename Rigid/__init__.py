"""Drop-in ``Rigid`` package: same import surface as the reference's install layout
(site-packages/Rigid/{__init__.py, Rigid.py, c_rigid*.so}; /root/reference/src/__init__.py:1,
CMakeLists.txt:24-27), served by the B200-native implementation in rigid_body_light_b200."""
from rigid_body_light_b200 import c_rigid  # noqa: F401  (``from Rigid import c_rigid`` works)
from rigid_body_light_b200.Rigid import RigidBody  # noqa: F401
