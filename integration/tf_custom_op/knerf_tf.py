"""Python side of the TensorFlow custom ops (knerf_tf_ops.cc): loads the op library, registers the gradient of the
compositing op and shows the drop-in replacement of `NeRFUtils.render_image_depth_chunk`
(keras_nerf/model/nerf/utils.py:16-58).  Needs TensorFlow >= 2.9 -- not importable in the build image, so this file
is documentation that compiles where TensorFlow exists; the tested shim is keras_nerf_b200/ (ctypes + DLPack)."""
import os

import tensorflow as tf

_ops = tf.load_op_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "knerf_tf_ops.so"))


@tf.RegisterGradient("KnerfCompositeForward")
def _composite_grad(op, d_image, d_depth, d_weights):
    """autodiff of utils.py:32-58 w.r.t. rgb and sigma for a loss that depends on the image only (the reference's
    losses do: nerf.py:372-373); depth / weights carry no gradient in train_step"""
    rgb, sigma, t = op.inputs
    rgbsigma = tf.concat([rgb, sigma], axis=-1)
    d = _ops.knerf_composite_backward(rgbsigma, t, d_image, white_background=op.get_attr("white_background"),
                                      clip=op.get_attr("clip"), epsilon=op.get_attr("epsilon"))
    return d[..., :3], d[..., 3:], None


def render_image_depth_chunk(self, rgb, sigma, sample_points, epsilon=1e-10):
    """drop-in body for NeRFUtils.render_image_depth_chunk"""
    return _ops.knerf_composite_forward(rgb, sigma, sample_points, white_background=self.white_background, clip=True,
                                        epsilon=epsilon)


def fine_points(coarse_points, coarse_weights, n_fine):
    """nerf.py:182-191 in one op: mid points, inverse-cdf samples (TF-GPU gather semantics), concat and sort"""
    u = tf.random.uniform([tf.shape(coarse_points)[0], n_fine])
    return _ops.knerf_sample_fine(coarse_points, tf.stop_gradient(coarse_weights), u, oob_mode=0)
