"""Import the UNMODIFIED reference modules with ROS stubbed out.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Used only by
``tests/golden/make_golden.py`` (to generate the committed fixtures) and by the
tests that re-check the oracle against the live reference when
``/root/reference`` happens to be present (it is not on the GPU box).

``auv_particle.py`` imports ``rospy`` (``:5``, never used in the file) and
``tf.transformations`` (``:8-10``); neither exists in this image.  We inject an
empty ``rospy`` module and ``oracle.tf_stub`` as ``tf.transformations`` into
``sys.modules`` and then import ``auv_particle`` and ``resampling`` from the
reference checkout as they lie -- no reference source is copied or edited.
``auv_pf.py`` itself cannot be imported (needs a ROS master, ``tf2_ros`` and
message packages, and spins inside its constructor, ``auv_pf.py:25-119``); its
``resample`` bookkeeping is restated in ``pf_oracle.resample_bookkeeping``.
"""
import importlib
import os
import sys
import types

DEFAULT_REF = "/root/reference/auv_particle_filter/scripts"


def available(path=DEFAULT_REF):
    return os.path.isfile(os.path.join(path, "auv_particle.py"))


def load(path=DEFAULT_REF):
    """Return ``(auv_particle, resampling)`` modules of the reference."""
    if not available(path):
        raise FileNotFoundError(path)
    from . import tf_stub
    if "rospy" not in sys.modules:
        sys.modules["rospy"] = types.ModuleType("rospy")
    if "tf" not in sys.modules:
        tf_mod = types.ModuleType("tf")
        tf_mod.transformations = tf_stub
        sys.modules["tf"] = tf_mod
        sys.modules["tf.transformations"] = tf_stub
    if path not in sys.path:
        sys.path.insert(0, path)
    auv_particle = importlib.import_module("auv_particle")
    resampling = importlib.import_module("resampling")
    return auv_particle, resampling


class Msg(object):
    """Attribute bag standing in for ROS messages (Odometry / PointStamped)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_odom(v, wz, quat_xyzw, z):
    """ROS ``nav_msgs/Odometry``-shaped object with the attributes that
    ``Particle.motion_pred`` reads (``auv_particle.py:43-70``)."""
    lin = Msg(x=float(v[0]), y=float(v[1]), z=float(v[2]))
    ang = Msg(x=0.0, y=0.0, z=float(wz))
    ori = Msg(x=float(quat_xyzw[0]), y=float(quat_xyzw[1]),
              z=float(quat_xyzw[2]), w=float(quat_xyzw[3]))
    pos = Msg(x=0.0, y=0.0, z=float(z))
    return Msg(twist=Msg(twist=Msg(linear=lin, angular=ang)),
               pose=Msg(pose=Msg(orientation=ori, position=pos)))


def make_point(x, y, z=0.0):
    """``geometry_msgs/PointStamped``-shaped object (``auv_particle.py:105``)."""
    return Msg(point=Msg(x=float(x), y=float(y), z=float(z)))
