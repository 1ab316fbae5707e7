// reagent_b200 -- stand-alone fused MLP forward over row tiles (inference / scoring).
// Used by the nn.Module.forward of the models in reagent_b200/models (the reference's
// FullyConnectedNetwork.forward, reagent/models/fully_connected_network.py:157-163, and
// the critic's cat(state, action) input, reagent/models/critic.py:76-92).
#include "rb200_rows.cuh"

namespace rb200 {

struct FwdDev {
  const float* in0; int d0;   // [B, d0]
  const float* in1; int d1;   // [B, d1] or nullptr (concatenated after in0)
  float* out;                 // [B, dims[L]]
  int batch, ld_in, ld_h, ld_o;
  rb200_net_ws_t ws;
  int save;
};

template <int NT, int TM, int KC>
__global__ void __launch_bounds__(NT, 1) mlp_fwd_rows_kernel(const Mlp net, const FwdDev p) {
  constexpr int R = (NT / 64) * TM;
  extern __shared__ __align__(16) float smem[];
  tile_smem_zero_all<NT>(smem);
  float* Wst = smem;
  float* xin = Wst + 2 * wstage_floats<KC>();
  float* hA = xin + R * p.ld_in;
  float* hB = hA + R * p.ld_h;
  float* xo = hB + R * p.ld_h;
  const int row0 = blockIdx.x * R;
  // cat(in0, in1): in0 occupies columns [0,d0), in1 columns [d0, d0+d1)
  if (p.in1 == nullptr) {
    tile_load_rows<NT, R>(xin, p.ld_in, p.in0, p.d0, p.d0, row0, p.batch);
  } else {
    const int D = p.d0 + p.d1, D4 = round_up4(D);
    for (int idx = threadIdx.x; idx < R * D4; idx += NT) {
      const int r = idx / D4, c = idx - r * D4;
      float v = 0.f;
      if (row0 + r < p.batch) {
        if (c < p.d0) v = p.in0[(size_t)(row0 + r) * p.d0 + c];
        else if (c < D) v = p.in1[(size_t)(row0 + r) * p.d1 + (c - p.d0)];
      }
      xin[r * p.ld_in + c] = v;
    }
  }
  __syncthreads();
  tile_mlp_fwd<NT, TM, KC>(net, xin, p.ld_in, hA, hB, p.ld_h, xo, p.ld_o, Wst,
                           p.save ? p.ws.hidden : nullptr, row0, p.batch);
  const int DO = net.dims[net.n_layers];
  tile_store_rows<NT, R>(xo, p.ld_o, p.out, DO, DO, row0, p.batch);
}

#define RB200_LAUNCH_FWD(NT_, TM_, KC_, grid, smem, stream, ...)                                   \
  do {                                                                                        \
    auto kfn = mlp_fwd_rows_kernel<NT_, TM_, KC_>;                                                 \
    static SmemOptIn optin_ = {};                                                             \
    {                                                                                         \
      cudaError_t e_ = ensure_dynamic_smem(kfn, optin_, (size_t)(smem));                      \
      if (e_ != cudaSuccess) return check_cuda(e_, "cudaFuncSetAttribute(mlp_fwd)");                                       \
    }                                                                                         \
    kfn<<<grid, NT_, smem, stream>>>(__VA_ARGS__);                                       \
  } while (0)

}  // namespace rb200

using namespace rb200;

extern "C" int rb200_mlp_forward(const rb200_mlp_t* net, const float* in0, int32_t d0,
                                 const float* in1, int32_t d1, int32_t batch, float* out,
                                 const rb200_net_ws_t* save_hidden, void* stream) {
  if (!net || !in0 || !out) { set_last_error("rb200_mlp_forward: null argument"); return RB200_E_INVALID; }
  if (int rc = validate_mlp(net, "net")) return rc;
  if (batch <= 0) { set_last_error("rb200_mlp_forward: batch must be positive"); return RB200_E_INVALID; }
  if (d0 + (in1 ? d1 : 0) != net->dims[0]) {
    set_last_error("rb200_mlp_forward: input width %d != dims[0]=%d", d0 + (in1 ? d1 : 0), net->dims[0]);
    return RB200_E_INVALID;
  }
  FwdDev p;
  p.in0 = in0; p.d0 = d0; p.in1 = in1; p.d1 = in1 ? d1 : 0; p.out = out; p.batch = batch;
  p.save = save_hidden ? 1 : 0;
  if (save_hidden) p.ws = *save_hidden; else p.ws = rb200_net_ws_t{};
  const int DO = net->dims[net->n_layers];
  const int hmax = mlp_max_hidden(net);
  p.ld_o = round_up4(DO) + 4;
  if (p.ld_o > 1024 + 4) { set_last_error("rb200_mlp_forward: output width %d too large for the row-tile kernel", DO); return RB200_E_SMEM; }
  RowsCfg cfg = pick_rows_cfg(batch, net->dims[0], hmax, 1, 2, p.ld_o, 0);
  if (cfg.tm == 0) { set_last_error("rb200_mlp_forward: tile does not fit in shared memory"); return RB200_E_SMEM; }
  p.ld_in = cfg.ld_in; p.ld_h = cfg.ld_h;
  const Mlp m = make_mlp(net);
  const int grid = ceil_div(batch, rows_per_tile(cfg));
  cudaStream_t st = (cudaStream_t)stream;
  RB200_DISPATCH_ROWS(cfg, RB200_LAUNCH_FWD, grid, cfg.smem_bytes, st, m, p);
  return check_cuda(cudaGetLastError(), "mlp_fwd_rows_kernel launch");
}
