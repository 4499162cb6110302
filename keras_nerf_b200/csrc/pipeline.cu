// a6/a9/a10/a11: fused MLP entry points and the per-chunk coarse+fine pipelines
// (keras_nerf/model/nerf/nerf.py:175-227 render, :351-421 train).  Host-side orchestration only: every
// stage is an asynchronous launch on the caller's stream; nothing here synchronises.
#include "common.cuh"
#include "mlp_fp32.cuh"
#include "mlp_tc.cuh"
#include "comm.cuh"

namespace knerf {
int launch_sum_scale(const float* x, int64_t n, float scale, float* out, int accumulate, cudaStream_t st);

struct ChunkWs {
  float *rgbsigma, *d_pre, *t_sorted, *weights, *image, *sqerr;
  char* mlp;
  int64_t mlp_bytes;
};

static int carve(void* ws, int64_t bytes, int64_t R, int S, bool training, ChunkWs* c) {
  char* base = (char*)ws;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += (b + 255) & ~(size_t)255; return base + o; };
  const size_t rows = (size_t)R * S;
  c->rgbsigma = (float*)take(rows * 16);
  c->d_pre = training ? (float*)take(rows * 16) : nullptr;
  c->t_sorted = (float*)take(rows * 4);
  c->weights = (float*)take(rows * 4);
  c->image = (float*)take((size_t)R * 12);
  c->sqerr = (float*)take((size_t)R * 4);
  c->mlp = base + off;
  c->mlp_bytes = bytes - (int64_t)off;
  if (ws == nullptr || c->mlp_bytes < 0)
    return fail(KNERF_ERR_WORKSPACE, "workspace too small for chunk scratch (%lld bytes)", (long long)bytes);
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(KNERF_ERR_INVALID, "workspace must be 256-byte aligned");
  return KNERF_OK;
}

// `precision` carries the per-call option bits of knerf.h (KNERF_TC_ORDERED, KNERF_BWD_*_ONLY)
static int mlp_forward_impl(const Model& m, const float* params, const void* packed, const float* o, const float* d,
                            const float* t, int64_t R, int S, int precision_flags, bool training, float* rgbsigma,
                            char* ws, int64_t ws_bytes, cudaStream_t st) {
  const int64_t rows = R * S;
  const int precision = precision_flags & KNERF_PRECISION_MASK;
  if (precision == KNERF_FP32 || precision == KNERF_FP32_TC) {
    const Fp32Plan p = make_fp32_plan(m, rows, training, precision == KNERF_FP32_TC);
    if ((int64_t)p.total > ws_bytes)
      return fail(KNERF_ERR_WORKSPACE, "knerf_mlp_forward: workspace %lld < %lld bytes", (long long)ws_bytes,
                  (long long)p.total);
    float* X0 = (float*)(ws + p.off_x0);
    float* DIR = (float*)(ws + p.off_dir);
    KN_TRY(knerf_encode_position_and_directions(o, d, t, R, S, m.cfg.pos_emb_xyz, m.cfg.pos_emb_dir, X0, p.ldx, DIR,
                                                p.ldd, st));
    return fp32_forward_core(m, params, X0, p.ldx, DIR, p.ldd, rows, ws, p, rgbsigma, 4, rgbsigma + 3, 4, st, training);
  }
  if (precision == KNERF_BF16) {
    if (packed == nullptr) return fail(KNERF_ERR_INVALID, "KNERF_BF16 needs packed weights (knerf_pack_weights)");
    return tc_forward(m, params, packed, o, d, t, R, S, training, (precision_flags & KNERF_TC_ORDERED) != 0,
                      (precision_flags & KNERF_REC_FP8) != 0, rgbsigma, ws, ws_bytes, st);
  }
  return fail(KNERF_ERR_INVALID, "unknown precision %d", precision);
}

static int mlp_backward_impl(const Model& m, const float* params, const void* packed, const float* d_pre, int64_t R,
                             int S, int precision_flags, float* grads, char* ws, int64_t ws_bytes, cudaStream_t st) {
  const int64_t rows = R * S;
  const int precision = precision_flags & KNERF_PRECISION_MASK;
  if (precision == KNERF_FP32 || precision == KNERF_FP32_TC) {
    const Fp32Plan p = make_fp32_plan(m, rows, true, precision == KNERF_FP32_TC);
    if ((int64_t)p.total > ws_bytes)
      return fail(KNERF_ERR_WORKSPACE, "knerf_mlp_backward: workspace %lld < %lld bytes", (long long)ws_bytes,
                  (long long)p.total);
    return fp32_backward_core(m, params, (const float*)(ws + p.off_x0), p.ldx, (const float*)(ws + p.off_dir), p.ldd,
                              d_pre, rows, ws, p, grads, st);
  }
  if (precision == KNERF_BF16) {
    const int parts = (precision_flags & KNERF_BWD_DGRAD_ONLY) ? 1 : (precision_flags & KNERF_BWD_WGRAD_ONLY) ? 2 : 3;
    return tc_backward(m, params, packed, d_pre, R, S, grads, ws, ws_bytes, parts, (precision_flags & KNERF_REC_FP8) != 0,
                       st);
  }
  return fail(KNERF_ERR_INVALID, "unknown precision %d", precision);
}

static int check_model_for_rays(const knerf_config* cfg, Model* m) {
  KN_TRY(build_model(cfg, m));
  KN_CHECK_ARG(m->dx == 3 + 6 * cfg->pos_emb_xyz && m->dd == 3 + 6 * cfg->pos_emb_dir,
               "ray entry points need dx/dd = 3+6L (got dx=%d dd=%d)", m->dx, m->dd);
  return KNERF_OK;
}

}  // namespace knerf

using namespace knerf;

extern "C" int64_t knerf_packed_weight_bytes(const knerf_config* cfg) {
  Model m;
  if (build_model(cfg, &m) != KNERF_OK) return -1;
  const int64_t n = tc_packed_weight_bytes(m);
  if (n < 0)
    fail(KNERF_ERR_UNSUPPORTED, "KNERF_BF16 implements models of dense_units <= 256 (even), up to 8 layers, at most one skip concat (not into the heads), L_xyz <= 10, L_dir <= 4 (got %d x %d, skip %d, L=%d,%d)",
         m.n_layers, m.U, m.cfg.skip_layer, m.cfg.pos_emb_xyz, m.cfg.pos_emb_dir);
  return n;
}

extern "C" int knerf_pack_weights(const knerf_config* cfg, const float* params, void* packed, void* stream) {
  Model m;
  KN_TRY(build_model(cfg, &m));
  KN_CHECK_ARG(params && packed, "knerf_pack_weights: null pointer");
  return tc_pack_weights(m, params, packed, (cudaStream_t)stream);
}

extern "C" int knerf_mlp_forward(const knerf_config* cfg, const float* params, const void* packed, const float* o,
                                 const float* d, const float* t, int64_t R, int S, int precision, int training,
                                 float* rgbsigma, void* workspace, int64_t workspace_bytes, void* stream) {
  Model m;
  KN_TRY(check_model_for_rays(cfg, &m));
  KN_CHECK_ARG(params && o && d && t && rgbsigma && workspace && R >= 0 && S > 0, "knerf_mlp_forward: bad arguments");
  if (R == 0) return KNERF_OK;
  return mlp_forward_impl(m, params, packed, o, d, t, R, S, precision, training != 0, rgbsigma, (char*)workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

extern "C" int knerf_mlp_backward(const knerf_config* cfg, const float* params, const void* packed,
                                  const float* d_pre, int64_t R, int S, int precision, float* grads, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  Model m;
  KN_TRY(check_model_for_rays(cfg, &m));
  KN_CHECK_ARG(params && d_pre && grads && workspace && R >= 0 && S > 0, "knerf_mlp_backward: bad arguments");
  if (R == 0) return KNERF_OK;
  return mlp_backward_impl(m, params, packed, d_pre, R, S, precision, grads, (char*)workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

extern "C" int knerf_render_chunk(const knerf_config* cfg, const float* params_coarse, const float* params_fine,
                                  const void* packed_coarse, const void* packed_fine, const float* o, const float* d,
                                  const float* t_coarse, int64_t R, const float* u_fine, uint64_t seed,
                                  int white_background, int oob_mode, int precision, float* image_c, float* depth_c,
                                  float* weights_c, float* image_f, float* depth_f, float* weights_f,
                                  float* t_fine_sorted, void* workspace, int64_t workspace_bytes, void* stream) {
  Model m;
  KN_TRY(check_model_for_rays(cfg, &m));
  KN_CHECK_ARG(params_coarse && params_fine && o && d && t_coarse && R >= 0, "knerf_render_chunk: null argument");
  if (R == 0) return KNERF_OK;
  const int Nc = cfg->n_coarse, Nf = cfg->n_fine, S = Nc + Nf;
  cudaStream_t st = (cudaStream_t)stream;
  ChunkWs c;
  KN_TRY(carve(workspace, workspace_bytes, R, S, false, &c));
  float* wc = weights_c ? weights_c : c.weights;
  float* ts = t_fine_sorted ? t_fine_sorted : c.t_sorted;
  // coarse pass (nerf.py:220-221)
  KN_TRY(mlp_forward_impl(m, params_coarse, packed_coarse, o, d, t_coarse, R, Nc, precision, false, c.rgbsigma, c.mlp,
                          c.mlp_bytes, st));
  KN_TRY(knerf_composite_forward(c.rgbsigma, nullptr, nullptr, t_coarse, R, Nc, white_background, 1, 1e-10f, image_c,
                                 depth_c, wc, nullptr, st));
  // hierarchical sampling + sort (nerf.py:182-191)
  KN_TRY(knerf_sample_fine(t_coarse, nullptr, wc, u_fine, seed, nullptr, R, Nc, Nf, oob_mode, ts, nullptr, nullptr,
                           nullptr, nullptr, st));
  // fine pass on all Nc+Nf depths (nerf.py:224-225)
  KN_TRY(mlp_forward_impl(m, params_fine, packed_fine, o, d, ts, R, S, precision, false, c.rgbsigma, c.mlp,
                          c.mlp_bytes, st));
  KN_TRY(knerf_composite_forward(c.rgbsigma, nullptr, nullptr, ts, R, S, white_background, 1, 1e-10f, image_f,
                                 depth_f, weights_f, nullptr, st));
  return KNERF_OK;
}

// one chunk of train_step; comm != nullptr: the all-reduce of each network's accumulated gradient is enqueued on
// comm_stream right behind that network's backward (knerf_train_chunk_dp)
static int train_chunk_impl(const knerf_config* cfg, const float* params_coarse, const float* params_fine,
                            const void* packed_coarse, const void* packed_fine, const float* o, const float* d,
                            const float* t_coarse, const float* target_rgb, int64_t R, const float* u_fine,
                            uint64_t seed, int white_background, int oob_mode, int precision, float grad_scale,
                            float* grads_coarse, float* grads_fine, float* losses, float* image_c, float* image_f,
                            void* workspace, int64_t workspace_bytes, void* stream, knerf_comm* comm,
                            void* comm_stream) {
  Model m;
  KN_TRY(check_model_for_rays(cfg, &m));
  KN_CHECK_ARG(params_coarse && params_fine && o && d && t_coarse && target_rgb && grads_coarse && grads_fine &&
                   losses && R > 0,
               "knerf_train_chunk: null argument");
  const int Nc = cfg->n_coarse, Nf = cfg->n_fine, S = Nc + Nf;
  cudaStream_t st = (cudaStream_t)stream;
  ChunkWs c;
  KN_TRY(carve(workspace, workspace_bytes, R, S, true, &c));
  // d/dC of mean_{R x 3}(C - target)^2, scaled by 1/sequential_chunks (nerf.py:372-373,383-384)
  const float loss_scale = grad_scale * 2.0f / (3.0f * (float)R);
  const float mse_scale = grad_scale / (3.0f * (float)R);

  // ---- coarse network (nerf.py:361-388) ----
  KN_TRY(mlp_forward_impl(m, params_coarse, packed_coarse, o, d, t_coarse, R, Nc, precision, true, c.rgbsigma, c.mlp,
                          c.mlp_bytes, st));
  KN_TRY(knerf_composite_forward(c.rgbsigma, nullptr, nullptr, t_coarse, R, Nc, white_background, 1, 1e-10f, image_c,
                                 nullptr, c.weights, nullptr, st));
  KN_TRY(knerf_composite_backward(c.rgbsigma, t_coarse, R, Nc, white_background, 1, 1e-10f, nullptr, target_rgb,
                                  loss_scale, 1, c.d_pre, c.sqerr, st));
  KN_TRY(launch_sum_scale(c.sqerr, R, mse_scale, losses, 1, st));
  KN_TRY(mlp_backward_impl(m, params_coarse, packed_coarse, c.d_pre, R, Nc, precision, grads_coarse, c.mlp,
                           c.mlp_bytes, st));
  // the coarse gradient is final: it crosses NVLink while the fine network runs
  if (comm != nullptr) KN_TRY(comm_allreduce_after(comm, grads_coarse, m.n_params, st, (cudaStream_t)comm_stream, 0));
  // ---- fine network; coarse weights are constants, no gradient through the sampler (nerf.py:390-417) ----
  KN_TRY(knerf_sample_fine(t_coarse, nullptr, c.weights, u_fine, seed, nullptr, R, Nc, Nf, oob_mode, c.t_sorted,
                           nullptr, nullptr, nullptr, nullptr, st));
  KN_TRY(mlp_forward_impl(m, params_fine, packed_fine, o, d, c.t_sorted, R, S, precision, true, c.rgbsigma, c.mlp,
                          c.mlp_bytes, st));
  if (image_f != nullptr)
    KN_TRY(knerf_composite_forward(c.rgbsigma, nullptr, nullptr, c.t_sorted, R, S, white_background, 1, 1e-10f,
                                   image_f, nullptr, nullptr, nullptr, st));
  KN_TRY(knerf_composite_backward(c.rgbsigma, c.t_sorted, R, S, white_background, 1, 1e-10f, nullptr, target_rgb,
                                  loss_scale, 1, c.d_pre, c.sqerr, st));
  KN_TRY(launch_sum_scale(c.sqerr, R, mse_scale, losses + 1, 1, st));
  KN_TRY(mlp_backward_impl(m, params_fine, packed_fine, c.d_pre, R, S, precision, grads_fine, c.mlp, c.mlp_bytes, st));
  if (comm != nullptr) {
    KN_TRY(comm_allreduce_after(comm, grads_fine, m.n_params, st, (cudaStream_t)comm_stream, 1));
    KN_TRY(comm_join(comm, (cudaStream_t)comm_stream, st));   // `stream` continues once both reductions are done
  }
  return KNERF_OK;
}

extern "C" int knerf_train_chunk(const knerf_config* cfg, const float* params_coarse, const float* params_fine,
                                 const void* packed_coarse, const void* packed_fine, const float* o, const float* d,
                                 const float* t_coarse, const float* target_rgb, int64_t R, const float* u_fine,
                                 uint64_t seed, int white_background, int oob_mode, int precision, float grad_scale,
                                 float* grads_coarse, float* grads_fine, float* losses, float* image_c,
                                 float* image_f, void* workspace, int64_t workspace_bytes, void* stream) {
  return train_chunk_impl(cfg, params_coarse, params_fine, packed_coarse, packed_fine, o, d, t_coarse, target_rgb, R,
                          u_fine, seed, white_background, oob_mode, precision, grad_scale, grads_coarse, grads_fine,
                          losses, image_c, image_f, workspace, workspace_bytes, stream, nullptr, nullptr);
}

extern "C" int knerf_train_chunk_dp(const knerf_config* cfg, const float* params_coarse, const float* params_fine,
                                    const void* packed_coarse, const void* packed_fine, const float* o, const float* d,
                                    const float* t_coarse, const float* target_rgb, int64_t R, const float* u_fine,
                                    uint64_t seed, int white_background, int oob_mode, int precision,
                                    float grad_scale, float* grads_coarse, float* grads_fine, float* losses,
                                    float* image_c, float* image_f, void* workspace, int64_t workspace_bytes,
                                    void* stream, knerf_comm* comm, void* comm_stream, int reduce) {
  const bool dp = comm != nullptr && reduce != 0;
  if (dp) KN_CHECK_ARG(comm_stream != stream, "knerf_train_chunk_dp: comm_stream must differ from stream");
  return train_chunk_impl(cfg, params_coarse, params_fine, packed_coarse, packed_fine, o, d, t_coarse, target_rgb, R,
                          u_fine, seed, white_background, oob_mode, precision, grad_scale, grads_coarse, grads_fine,
                          losses, image_c, image_f, workspace, workspace_bytes, stream, dp ? comm : nullptr,
                          dp ? comm_stream : nullptr);
}
