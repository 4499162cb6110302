// TensorFlow custom ops over the C ABI of libknerf.so (include/knerf.h) -- the "thin TF custom-op shim" of the north
// star for hosts that stay in TensorFlow (the reference, naufalso/keras_nerf, is Python on TF >= 2.9).
//
// NOT PART OF THE DEFAULT BUILD: TensorFlow's headers are not in this image (no network), so this file is compiled
// only where they exist:
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++17 -shared -fPIC knerf_tf_ops.cc -o knerf_tf_ops.so -I../../include $TF_CFLAGS $TF_LFLAGS \
//       -DGOOGLE_CUDA=1 -L../../keras_nerf_b200/lib -lknerf -Wl,-rpath,'$ORIGIN/../../keras_nerf_b200/lib'
//
// and loaded with tf.load_op_library (knerf_tf.py next to this file registers the gradients).  Every op forwards
// to ONE entry point of knerf.h with the tensors' device pointers and TF's own GPU stream; nothing is copied and no
// kernel lives here.  Op <-> reference function:
//
//   KnerfCompositeForward   NeRFUtils.render_image_depth_chunk          keras_nerf/model/nerf/utils.py:16-58
//   KnerfCompositeBackward  (its gradient, registered in knerf_tf.py)   autodiff of nerf.py:361-377
//   KnerfSampleFine         fine_hierarchical_sampling_chunk + tf.sort  utils.py:60-97, nerf.py:182-191
//   KnerfGenerateRays       RaysGenerator.__call__                      keras_nerf/data/rays.py:69-130
//   KnerfRenderChunk        NeRF.predict_and_render_chunk               nerf.py:218-227
//   KnerfTrainChunk         one iteration of NeRF.train_step's loop     nerf.py:351-421
//   KnerfAdamStep           optimizer.apply_gradients                   nerf.py:455-458
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include "knerf.h"

namespace tf = tensorflow;
using GPUDevice = Eigen::GpuDevice;

namespace {

inline void* tf_stream(tf::OpKernelContext* ctx) { return (void*)ctx->eigen_device<GPUDevice>().stream(); }

inline const float* in_f(tf::OpKernelContext* ctx, int i) { return ctx->input(i).flat<float>().data(); }

#define KNERF_TF_CALL(ctx, expr)                                                                        \
  do {                                                                                                  \
    const int _rc = (expr);                                                                             \
    OP_REQUIRES(ctx, _rc == 0, tf::errors::Internal("libknerf: ", knerf_last_error(), " (", _rc, ")")); \
  } while (0)

knerf_config config_from_attrs(tf::OpKernelConstruction* c) {
  knerf_config cfg{};
  int v = 0;
  auto get = [&](const char* name, int32_t* dst) { if (c->GetAttr(name, &v).ok()) *dst = v; };
  get("n_coarse", &cfg.n_coarse); get("n_fine", &cfg.n_fine);
  get("pos_emb_xyz", &cfg.pos_emb_xyz); get("pos_emb_dir", &cfg.pos_emb_dir);
  get("n_layers", &cfg.n_layers); get("dense_units", &cfg.dense_units); get("skip_layer", &cfg.skip_layer);
  return cfg;   // dx = dd = 0: 3 + 6 L
}

}  // namespace

// ---- a7: render_image_depth_chunk -----------------------------------------------------------------------------
REGISTER_OP("KnerfCompositeForward")
    .Input("rgb: float").Input("sigma: float").Input("sample_points: float")
    .Attr("white_background: bool = false").Attr("clip: bool = true").Attr("epsilon: float = 1e-10")
    .Output("image: float").Output("depth: float").Output("weights: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      tf::shape_inference::ShapeHandle t = c->input(2);
      c->set_output(0, c->Matrix(c->Dim(t, 0), 3));
      c->set_output(1, c->Vector(c->Dim(t, 0)));
      c->set_output(2, t);
      return tf::OkStatus();
    });

class KnerfCompositeForwardOp : public tf::OpKernel {
 public:
  explicit KnerfCompositeForwardOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("white_background", &white_));
    OP_REQUIRES_OK(c, c->GetAttr("clip", &clip_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& t = ctx->input(2);
    const int64_t R = t.dim_size(0);
    const int S = (int)t.dim_size(1);
    tf::Tensor *image, *depth, *weights;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {R, 3}, &image));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {R}, &depth));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, {R, S}, &weights));
    KNERF_TF_CALL(ctx, knerf_composite_forward(nullptr, in_f(ctx, 0), in_f(ctx, 1), in_f(ctx, 2), R, S, white_, clip_,
                                               eps_, image->flat<float>().data(), depth->flat<float>().data(),
                                               weights->flat<float>().data(), nullptr, tf_stream(ctx)));
  }
 private:
  bool white_, clip_;
  float eps_;
};
REGISTER_KERNEL_BUILDER(Name("KnerfCompositeForward").Device(tf::DEVICE_GPU), KnerfCompositeForwardOp);

// gradient of the above w.r.t. (rgb, sigma) given dL/dimage; rgbsigma packed [R,S,4] by the Python wrapper
REGISTER_OP("KnerfCompositeBackward")
    .Input("rgbsigma: float").Input("sample_points: float").Input("dimage: float")
    .Attr("white_background: bool = false").Attr("clip: bool = true").Attr("epsilon: float = 1e-10")
    .Output("d_rgbsigma: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) { c->set_output(0, c->input(0)); return tf::OkStatus(); });

class KnerfCompositeBackwardOp : public tf::OpKernel {
 public:
  explicit KnerfCompositeBackwardOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("white_background", &white_));
    OP_REQUIRES_OK(c, c->GetAttr("clip", &clip_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& t = ctx->input(1);
    tf::Tensor* out;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, ctx->input(0).shape(), &out));
    KNERF_TF_CALL(ctx, knerf_composite_backward(in_f(ctx, 0), in_f(ctx, 1), t.dim_size(0), (int)t.dim_size(1), white_,
                                                clip_, eps_, in_f(ctx, 2), nullptr, 1.0f, /*through_activations=*/0,
                                                out->flat<float>().data(), nullptr, tf_stream(ctx)));
  }
 private:
  bool white_, clip_;
  float eps_;
};
REGISTER_KERNEL_BUILDER(Name("KnerfCompositeBackward").Device(tf::DEVICE_GPU), KnerfCompositeBackwardOp);

// ---- a8 + the sort of a9 -----------------------------------------------------------------------------------------
REGISTER_OP("KnerfSampleFine")
    .Input("t_coarse: float").Input("weights: float").Input("u: float")
    .Attr("oob_mode: int = 0")
    .Output("t_sorted: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      tf::shape_inference::DimensionHandle s;
      TF_RETURN_IF_ERROR(c->Add(c->Dim(c->input(0), 1), c->Dim(c->input(2), 1), &s));
      c->set_output(0, c->Matrix(c->Dim(c->input(0), 0), s));
      return tf::OkStatus();
    });

class KnerfSampleFineOp : public tf::OpKernel {
 public:
  explicit KnerfSampleFineOp(tf::OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("oob_mode", &oob_)); }
  void Compute(tf::OpKernelContext* ctx) override {
    const int64_t R = ctx->input(0).dim_size(0);
    const int Nc = (int)ctx->input(0).dim_size(1), Nf = (int)ctx->input(2).dim_size(1);
    tf::Tensor* out;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {R, Nc + Nf}, &out));
    KNERF_TF_CALL(ctx, knerf_sample_fine(in_f(ctx, 0), nullptr, in_f(ctx, 1), in_f(ctx, 2), 0, nullptr, R, Nc, Nf, oob_,
                                         out->flat<float>().data(), nullptr, nullptr, nullptr, nullptr, tf_stream(ctx)));
  }
 private:
  int oob_;
};
REGISTER_KERNEL_BUILDER(Name("KnerfSampleFine").Device(tf::DEVICE_GPU), KnerfSampleFineOp);

// ---- a3 ----------------------------------------------------------------------------------------------------------
REGISTER_OP("KnerfGenerateRays")
    .Input("camera_params: float")            // [4,4], HOST memory (the pose is read on the host)
    .Attr("height: int").Attr("width: int").Attr("focal: float").Attr("near: float").Attr("far: float")
    .Attr("n_sample: int").Attr("seed: int = 0")
    .Output("ray_origin: float").Output("ray_direction: float").Output("sample_points: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      int h, w, n;
      TF_RETURN_IF_ERROR(c->GetAttr("height", &h));
      TF_RETURN_IF_ERROR(c->GetAttr("width", &w));
      TF_RETURN_IF_ERROR(c->GetAttr("n_sample", &n));
      c->set_output(0, c->MakeShape({h, w, 3}));
      c->set_output(1, c->MakeShape({h, w, 3}));
      c->set_output(2, c->MakeShape({h, w, n}));
      return tf::OkStatus();
    });

class KnerfGenerateRaysOp : public tf::OpKernel {
 public:
  explicit KnerfGenerateRaysOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("height", &h_)); OP_REQUIRES_OK(c, c->GetAttr("width", &w_));
    OP_REQUIRES_OK(c, c->GetAttr("focal", &focal_)); OP_REQUIRES_OK(c, c->GetAttr("near", &near_));
    OP_REQUIRES_OK(c, c->GetAttr("far", &far_)); OP_REQUIRES_OK(c, c->GetAttr("n_sample", &n_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    tf::Tensor *o, *d, *t;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {h_, w_, 3}, &o));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {h_, w_, 3}, &d));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, {h_, w_, n_}, &t));
    // every call draws fresh stratified jitter, like tf.random.uniform (rays.py:120): seed + call counter
    const uint64_t seed = (uint64_t)seed_ * 0x9E3779B97F4A7C15ull + calls_++;
    KNERF_TF_CALL(ctx, knerf_generate_rays(in_f(ctx, 0), h_, w_, focal_, near_, far_, n_, nullptr, seed,
                                           o->flat<float>().data(), d->flat<float>().data(), t->flat<float>().data(),
                                           tf_stream(ctx)));
  }
 private:
  int h_, w_, n_, seed_;
  float focal_, near_, far_;
  std::atomic<uint64_t> calls_{0};
};
REGISTER_KERNEL_BUILDER(Name("KnerfGenerateRays").Device(tf::DEVICE_GPU).HostMemory("camera_params"), KnerfGenerateRaysOp);

// ---- a9/a10: coarse + fine pass of one chunk -------------------------------------------------------------------
// (`precision` goes to the C ABI as it is: a knerf_precision value with the per-call option bits of knerf.h OR-ed in,
//  e.g. KNERF_BF16 | KNERF_REC_FP8 = 0x801 for the tensor-core mode with fp8 records between its training kernels)
#define KNERF_MODEL_ATTRS                                                                                        \
  .Attr("n_coarse: int = 64").Attr("n_fine: int = 128").Attr("pos_emb_xyz: int = 10").Attr("pos_emb_dir: int = 4") \
  .Attr("n_layers: int = 8").Attr("dense_units: int = 256").Attr("skip_layer: int = 4")                           \
  .Attr("white_background: bool = false").Attr("precision: int = 0").Attr("oob_mode: int = 0")

REGISTER_OP("KnerfRenderChunk")
    .Input("params_coarse: float").Input("params_fine: float")   // flat Keras-order buffers (knerf_layer_table)
    .Input("packed_coarse: uint8").Input("packed_fine: uint8")   // knerf_pack_weights output, or empty for fp32
    .Input("ray_origin: float").Input("ray_direction: float").Input("coarse_points: float").Input("u_fine: float")
    KNERF_MODEL_ATTRS
    .Output("image_coarse: float").Output("depth_coarse: float").Output("weights_coarse: float")
    .Output("image_fine: float").Output("depth_fine: float").Output("weights_fine: float")
    .SetShapeFn(tf::shape_inference::UnknownShape);

class KnerfRenderChunkOp : public tf::OpKernel {
 public:
  explicit KnerfRenderChunkOp(tf::OpKernelConstruction* c) : OpKernel(c), cfg_(config_from_attrs(c)) {
    OP_REQUIRES_OK(c, c->GetAttr("white_background", &white_));
    OP_REQUIRES_OK(c, c->GetAttr("precision", &prec_));
    OP_REQUIRES_OK(c, c->GetAttr("oob_mode", &oob_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const int64_t R = ctx->input(4).dim_size(0);
    const int Nc = cfg_.n_coarse, S = cfg_.n_coarse + cfg_.n_fine;
    tf::Tensor* out[6];
    const tf::TensorShape shapes[6] = {{R, 3}, {R}, {R, Nc}, {R, 3}, {R}, {R, S}};
    for (int i = 0; i < 6; ++i) OP_REQUIRES_OK(ctx, ctx->allocate_output(i, shapes[i], &out[i]));
    const int64_t ws_bytes = knerf_workspace_bytes(&cfg_, R * S, prec_ & KNERF_PRECISION_MASK, 0);
    OP_REQUIRES(ctx, ws_bytes >= 0, tf::errors::InvalidArgument("libknerf: ", knerf_last_error()));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, {ws_bytes + 256}, &ws));
    char* wsp = (char*)(((uintptr_t)ws.flat<uint8_t>().data() + 255) & ~(uintptr_t)255);
    auto packed = [&](int i) -> const void* {
      return ctx->input(i).NumElements() ? (const void*)ctx->input(i).flat<uint8_t>().data() : nullptr;
    };
    KNERF_TF_CALL(ctx, knerf_render_chunk(&cfg_, in_f(ctx, 0), in_f(ctx, 1), packed(2), packed(3), in_f(ctx, 4),
                                          in_f(ctx, 5), in_f(ctx, 6), R, in_f(ctx, 7), 0, white_, oob_, prec_,
                                          out[0]->flat<float>().data(), out[1]->flat<float>().data(),
                                          out[2]->flat<float>().data(), out[3]->flat<float>().data(),
                                          out[4]->flat<float>().data(), out[5]->flat<float>().data(), nullptr, wsp,
                                          ws_bytes, tf_stream(ctx)));
  }
 private:
  knerf_config cfg_;
  bool white_;
  int prec_, oob_;
};
REGISTER_KERNEL_BUILDER(Name("KnerfRenderChunk").Device(tf::DEVICE_GPU), KnerfRenderChunkOp);

// ---- a11/a12: one chunk of train_step; gradients are ACCUMULATED into resource-like buffers passed by reference ----
REGISTER_OP("KnerfTrainChunk")
    .Input("params_coarse: float").Input("params_fine: float").Input("packed_coarse: uint8").Input("packed_fine: uint8")
    .Input("ray_origin: float").Input("ray_direction: float").Input("coarse_points: float").Input("target_rgb: float")
    .Input("u_fine: float").Input("grads_coarse: Ref(float)").Input("grads_fine: Ref(float)").Input("losses: Ref(float)")
    KNERF_MODEL_ATTRS.Attr("grad_scale: float = 1.0")
    .Output("image_coarse: float").Output("image_fine: float")
    .SetShapeFn(tf::shape_inference::UnknownShape);

class KnerfTrainChunkOp : public tf::OpKernel {
 public:
  explicit KnerfTrainChunkOp(tf::OpKernelConstruction* c) : OpKernel(c), cfg_(config_from_attrs(c)) {
    OP_REQUIRES_OK(c, c->GetAttr("white_background", &white_));
    OP_REQUIRES_OK(c, c->GetAttr("precision", &prec_));
    OP_REQUIRES_OK(c, c->GetAttr("oob_mode", &oob_));
    OP_REQUIRES_OK(c, c->GetAttr("grad_scale", &scale_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const int64_t R = ctx->input(4).dim_size(0);
    const int S = cfg_.n_coarse + cfg_.n_fine;
    tf::Tensor *ic, *fi;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {R, 3}, &ic));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {R, 3}, &fi));
    const int64_t ws_bytes = knerf_workspace_bytes(&cfg_, R * S, prec_ & KNERF_PRECISION_MASK, 1);
    OP_REQUIRES(ctx, ws_bytes >= 0, tf::errors::InvalidArgument("libknerf: ", knerf_last_error()));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, {ws_bytes + 256}, &ws));
    char* wsp = (char*)(((uintptr_t)ws.flat<uint8_t>().data() + 255) & ~(uintptr_t)255);
    auto packed = [&](int i) -> const void* {
      return ctx->input(i).NumElements() ? (const void*)ctx->input(i).flat<uint8_t>().data() : nullptr;
    };
    tf::Tensor gc = ctx->mutable_input(9, true), gf = ctx->mutable_input(10, true), ls = ctx->mutable_input(11, true);
    KNERF_TF_CALL(ctx, knerf_train_chunk(&cfg_, in_f(ctx, 0), in_f(ctx, 1), packed(2), packed(3), in_f(ctx, 4),
                                         in_f(ctx, 5), in_f(ctx, 6), in_f(ctx, 7), R, in_f(ctx, 8), 0, white_, oob_,
                                         prec_, scale_, gc.flat<float>().data(), gf.flat<float>().data(),
                                         ls.flat<float>().data(), ic->flat<float>().data(), fi->flat<float>().data(),
                                         wsp, ws_bytes, tf_stream(ctx)));
  }
 private:
  knerf_config cfg_;
  bool white_;
  int prec_, oob_;
  float scale_;
};
REGISTER_KERNEL_BUILDER(Name("KnerfTrainChunk").Device(tf::DEVICE_GPU), KnerfTrainChunkOp);

// ---- a13: Keras Adam on the flat buffers ---------------------------------------------------------------------------
REGISTER_OP("KnerfAdamStep")
    .Input("params: Ref(float)").Input("grads: Ref(float)").Input("m: Ref(float)").Input("v: Ref(float)")
    .Input("step: int64")                      // HOST scalar, 1-based
    .Attr("learning_rate: float = 1e-3").Attr("beta_1: float = 0.9").Attr("beta_2: float = 0.999")
    .Attr("epsilon: float = 1e-7").Attr("zero_grads: bool = true")
    .SetShapeFn(tf::shape_inference::NoOutputs);

class KnerfAdamStepOp : public tf::OpKernel {
 public:
  explicit KnerfAdamStepOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("learning_rate", &lr_)); OP_REQUIRES_OK(c, c->GetAttr("beta_1", &b1_));
    OP_REQUIRES_OK(c, c->GetAttr("beta_2", &b2_)); OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("zero_grads", &zero_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    tf::Tensor p = ctx->mutable_input(0, true), g = ctx->mutable_input(1, true), m = ctx->mutable_input(2, true),
               v = ctx->mutable_input(3, true);
    const int64_t step = ctx->input(4).scalar<int64_t>()();
    KNERF_TF_CALL(ctx, knerf_adam_step(p.flat<float>().data(), g.flat<float>().data(), m.flat<float>().data(),
                                       v.flat<float>().data(), p.NumElements(), lr_, b1_, b2_, eps_, step, zero_,
                                       tf_stream(ctx)));
  }
 private:
  float lr_, b1_, b2_, eps_;
  bool zero_;
};
REGISTER_KERNEL_BUILDER(Name("KnerfAdamStep").Device(tf::DEVICE_GPU).HostMemory("step"), KnerfAdamStepOp);
