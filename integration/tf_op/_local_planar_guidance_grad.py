"""Gradient registration for the re-exposed TF op (counterpart of the reference's
tensorflow/custom_layer/_local_planar_guidance_grad.py:20-33): the `LocalPlanarGuidance` gradient is the
`LocalPlanarGuidanceGrad` op of the same library (integration/tf_op/bts_lpg_tf_op.cc over the bts_b200 C ABI).
Needs TensorFlow, which this image does not have -- shipped as the binding a maintainer would drop in."""
import os

import tensorflow as tf
from tensorflow.python.framework import ops

lpg = tf.load_op_library(os.environ.get("BTS_LPG_TF_LIB", "custom_layer/build/liblpg.so"))


@ops.RegisterGradient("LocalPlanarGuidance")
def _lpg_grad(op, depth_grad):
    return lpg.local_planar_guidance_grad(depth_grad, op.inputs[0], op.inputs[1])
