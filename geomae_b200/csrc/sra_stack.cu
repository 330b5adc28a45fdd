// Host-side executor of a stack of SRA EncoderLayers (forward and hand-derived backward).
// One C-ABI call runs every kernel of every layer of the stack on the given stream, so the Python
// host pays one call per stack and direction instead of one per kernel (the v1 path was CPU-bound).
// Layer math: models/sst/sst_basic_block.py:26-61,85-102 (post-norm EncoderLayer with window attention).
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace {

// Optional per-kernel-family device timing (bench.py's roofline leg): CUDA events recorded on the launching
// stream around every kernel of a family.  Off by default.
struct Prof {
  static constexpr int FAMILIES = 5;   // 0 tc_linear, 1 tc_wgrad, 2 attention fwd, 3 attention bwd, 4 layernorm bwd
  static constexpr int MAX_SPANS = 8192;
  bool on = false;
  int n = 0;
  cudaEvent_t start[MAX_SPANS], stop[MAX_SPANS];
  int family[MAX_SPANS];
  double flops[MAX_SPANS];
  double bytes[MAX_SPANS];
  int created = 0;
  cudaEvent_t* begin(int fam, double fl, double by, cudaStream_t st) {
    if (!on || n >= MAX_SPANS) return nullptr;
    if (n >= created) {
      if (cudaEventCreate(&start[n]) != cudaSuccess || cudaEventCreate(&stop[n]) != cudaSuccess) return nullptr;
      created = n + 1;
    }
    family[n] = fam;
    flops[n] = fl;
    bytes[n] = by;
    cudaEventRecord(start[n], st);
    return &stop[n++];
  }
};
Prof g_prof;

struct Span {
  cudaEvent_t* stop;
  cudaStream_t st;
  Span(int fam, double flops, void* stream, double bytes = 0.0) : st((cudaStream_t)stream) {
    stop = g_prof.begin(fam, flops, bytes, st);
  }
  ~Span() { if (stop) cudaEventRecord(*stop, st); }
};

int lin(const float* A, int lda, int n, int K, const float* W, int ldw, int w_rows, int mn, const float* bias, int N,
        float* out, int ldo, int prec, void* stream, geomae_linear_args* extra = nullptr, void* const* packed = nullptr) {
  geomae_linear_args a = extra ? *extra : geomae_linear_args{};
  a.A = A; a.lda = lda; a.n_rows = n; a.K = K; a.W = W; a.ldw = ldw; a.w_rows = w_rows; a.w_mn_major = mn;
  if (packed) { a.Wp_hi = packed[0]; a.Wp_lo = packed[1]; }
  a.bias = bias; a.N_total = N; a.out = out; a.ldo = ldo; a.precision = prec;
  // algorithmic bytes of the launch: A rows, bf16 weight image, output rows, plus every per-row side operand / output
  double bytes = (a.a_bf16 ? 2.0 : 4.0) * n * (double)K + (a.out_bf16 ? 2.0 : 4.0) * n * (double)N + 2.0 * (double)K * N;
  if (a.add_src) bytes += 4.0 * n * N;
  if (a.epilogue == 1) bytes += 4.0 * n * (N + 2.0);          // saved pre-LN rows + (mean, rstd)
  if (a.epilogue == 2) bytes += 4.0 * n * N;                  // GELU input
  if (a.epilogue == 3) bytes += 4.0 * n * (N + 2.0);          // saved pre-LN rows + (mean, rstd) read back
  if (a.dot_src) bytes += 4.0 * n * (N + 8.0);
  Span span(0, 2.0 * n * (double)K * N, stream, bytes);
  return geomae_tc_linear(&a, stream);
}

int wgrad(const float* dY, int ldy, const float* X, int ldx, int n, float* dW, int ldw, float* db, int M, int N,
          int prec, void* stream, const float* pos = nullptr, const int32_t* cell = nullptr, int pos_slabs = 0,
          int gelu = 0, int dy_bf16 = 0) {
  geomae_wgrad_args a{};
  a.dy_bf16 = dy_bf16;
  a.dY = dY; a.ldy = ldy; a.X = X; a.ldx = ldx; a.n_rows = n; a.pos_table = pos; a.tok_cell = cell;
  a.pos_slabs = pos_slabs; a.x_gelu = gelu; a.dW = dW; a.ldw = ldw; a.db = db; a.M_total = M; a.N_total = N;
  a.precision = prec;
  Span span(1, 2.0 * n * (double)M * N, stream, (dy_bf16 ? 2.0 : 4.0) * n * (double)M + 4.0 * n * (double)N + 4.0 * (double)M * N);
  return geomae_tc_wgrad(&a, stream);
}

#define GM_TRY(call)            \
  do {                          \
    int rc__ = (call);          \
    if (rc__) return rc__;      \
  } while (0)

}  // namespace

// ---- bf16 mode: two launches per layer.  The stack prologue projects x_in into layer 0's q|k|v; then per layer the
// window attention (sra_attention_tc.cu, bf16 in / bf16 out) and ONE chain kernel (sra_chain.cu) that runs out-proj + LN1
// -> FFN -> LN2 and already the in-projection of the next layer.
static int fused_forward(const geomae_sra_ctx* c, int32_t n_layers, const geomae_sra_layer* layers, const geomae_sra_saved* saved,
                  const float* x_in, void* stream) {
  const int64_t n = c->n_tokens;
  const double d = c->d_model, f = c->ffn;
  for (int l = 0; l < n_layers; ++l) GM_REQUIRE(saved[l].g, "sra_stack_forward: the bf16 path needs the g buffer of layer %d", l);
  GM_REQUIRE(saved[0].xb && c->pos16[0] && (c->pos16[1] || !c->shift[1].tok_cell),
             "sra_stack_forward: the bf16 path needs saved[0].xb and the pos16 scratch of every shift");
  for (int sft = 0; sft < 2; ++sft)      // gathered bf16 position rows per shift: operands of the in-projection weight gradients
    if (c->shift[sft].tok_cell) GM_TRY(geomae_pos_rows_bf16(c->pos_table, c->shift[sft].tok_cell, n, c->pos16[sft], stream));
  auto next_of = [&](geomae_chain_fwd_args& a, int l) {     // in-projection of layer l appended to the kernel
    const geomae_sra_layer& N = layers[l];
    a.p_in_proj_next = N.p_in_proj[0]; a.in_proj_b_next = N.in_proj_b;
    a.pos_table = c->pos_table; a.tok_cell_next = c->shift[N.shift].tok_cell;
    a.qkv16_next = saved[l].qkv;
  };
  const double in_flops = 2.0 * n * d * 3.0 * d, in_bytes = n * (2.0 * 3.0 * d) + 2.0 * 3.0 * d * d;
  {
    geomae_chain_fwd_args a{};
    a.n_tokens = n; a.mode = 2; a.x = x_in; a.xb16 = saved[0].xb;
    next_of(a, 0);
    Span span(0, in_flops, stream, 4.0 * n * d + in_bytes);
    GM_TRY(geomae_sra_chain_fwd(&a, stream));
  }
  const float* x = x_in;
  for (int l = 0; l < n_layers; ++l) {
    const geomae_sra_layer& L = layers[l];
    const geomae_sra_saved& S = saved[l];
    const geomae_sra_windows& w = c->shift[L.shift];
    {
      Span span(2, 0.0, stream, n * (2.0 * 3.0 * d + 2.0 * d + 4.0 * c->n_heads));
      GM_TRY(geomae_sra_attention_tc_fwd(S.qkv, n, c->n_heads, w.win_ptr, w.win_tok, w.tok_win, S.attn, S.lse, 1 | 8, stream));
    }
    geomae_chain_fwd_args a{};
    const bool has_next = l + 1 < n_layers;
    a.n_tokens = n; a.mode = has_next ? 3 : 1; a.x = x; a.attn = S.attn;
    a.p_out_proj = L.p_out_proj[0]; a.p_lin1 = L.p_lin1[0]; a.p_lin2 = L.p_lin2[0];
    a.out_proj_b = L.out_proj_b; a.lin1_b = L.lin1_b; a.lin2_b = L.lin2_b;
    a.norm1_w = L.norm1_w; a.norm1_b = L.norm1_b; a.norm2_w = L.norm2_w; a.norm2_b = L.norm2_b; a.ln_eps = L.ln_eps;
    a.xh1_16 = S.s1; a.st1 = S.st1; a.xh2_16 = S.s2; a.st2 = S.st2; a.z = S.z; a.u16 = S.u; a.g16 = S.g;
    if (has_next) next_of(a, l + 1);
    // algorithmic bytes: x + attn in; z (+ statistics), bf16 xhat1, xhat2, u, g out; bf16 weight images
    const double bytes = n * (4.0 * d + 2.0 * d + 4.0 * d + 16.0 + 2.0 * 2.0 * d + 2.0 * 2.0 * f) + 2.0 * (d * d + 2.0 * d * f) +
                         (has_next ? in_bytes : 0.0);
    Span span(0, 2.0 * n * (d * d + 2.0 * d * f) + (has_next ? in_flops : 0.0), stream, bytes);
    GM_TRY(geomae_sra_chain_fwd(&a, stream));
    x = S.z;
  }
  return GEOMAE_OK;
}

extern "C" int geomae_sra_stack_forward(const geomae_sra_ctx* c, int32_t n_layers, const geomae_sra_layer* layers,
                                        const geomae_sra_saved* saved, const float* x_in, void* stream) {
  GM_REQUIRE(c && layers && saved && (x_in || c->n_tokens == 0), "sra_stack_forward: null argument");
  GM_REQUIRE(c->d_model == 128 && c->n_heads == 8, "sra_stack: specialised for d_model 128 / 8 heads (got %d / %d)",
             c->d_model, c->n_heads);
  GM_REQUIRE(c->ffn % 128 == 0, "sra_stack: ffn width must be a multiple of 128");
  const int n = (int)c->n_tokens, d = c->d_model, f = c->ffn, p = c->precision;
  const int b16 = p == 1 ? 1 : 0;     // bf16 mode: tensors that are only MMA operands (qkv, du, da, dqkv) live in bf16
  if (n == 0) return GEOMAE_OK;
  const float* x = x_in;
  {  // refresh the packed bf16 weight images of every layer (weights change every optimiser step): one launch
    const float* W[64]; int32_t rows[64], cols[64]; void* hi[64]; void* lo[64];
    int k = 0;
    for (int l = 0; l < n_layers; ++l) {
      const geomae_sra_layer& L = layers[l];
      const float* ws[4] = {L.in_proj_w, L.out_proj_w, L.lin1_w, L.lin2_w};
      void* const* ps[4] = {L.p_in_proj, L.p_out_proj, L.p_lin1, L.p_lin2};
      const int r[4] = {3 * d, d, f, d}, cc[4] = {d, d, d, f};
      for (int i = 0; i < 4; ++i) {
        GM_REQUIRE(ps[i][0] && ps[i][1], "sra_stack_forward: layer %d has no packed-weight scratch", l);
        W[k] = ws[i]; rows[k] = r[i]; cols[k] = cc[i]; hi[k] = ps[i][0]; lo[k] = ps[i][1];
        if (++k == 64) { GM_TRY(geomae_pack_weights(k, W, rows, cols, hi, p == 1 ? nullptr : lo, stream)); k = 0; }
      }
    }
    if (k) GM_TRY(geomae_pack_weights(k, W, rows, cols, hi, p == 1 ? nullptr : lo, stream));
  }
  if (p == 1) return fused_forward(c, n_layers, layers, saved, x_in, stream);
  struct StableGuard { ~StableGuard() { gm_set_weights_stable(false); } } guard;
  gm_set_weights_stable(false);     // the first dense kernel directly follows the packing launch: no early weight fetch
  for (int l = 0; l < n_layers; ++l) {
    const geomae_sra_layer& L = layers[l];
    const geomae_sra_saved& S = saved[l];
    const geomae_sra_windows& w = c->shift[L.shift];
    geomae_linear_args e{};
    e.pos_table = c->pos_table; e.tok_cell = w.tok_cell; e.pos_slabs = 2;
    e.out_bf16 = b16;                 // q|k|v only ever feed the attention MMAs: stored as bf16 in the bf16 mode
    GM_TRY(lin(x, d, n, d, L.in_proj_w, d, 3 * d, 0, L.in_proj_b, 3 * d, S.qkv, 3 * d, p, stream, &e, L.p_in_proj));
    gm_set_weights_stable(true);    // from here on the images were complete before the predecessor kernel started
    {
      Span span(2, 0.0, stream, 4.0 * n * (3.0 * d + d + c->n_heads));
      if (p == 1)    // bf16 mode: QK^T / PV tiles on the tensor cores (sra_attention_tc.cu)
        GM_TRY(geomae_sra_attention_tc_fwd(S.qkv, n, c->n_heads, w.win_ptr, w.win_tok, w.tok_win, S.attn, S.lse, b16, stream));
      else
        GM_TRY(geomae_sra_attention_fwd(S.qkv, n, c->n_heads, w.win_ptr, w.win_tok, w.tok_win, S.attn, S.lse, stream));
    }
    geomae_linear_args e1{};
    e1.add_src = x; e1.ld_add = d; e1.ln_gamma = L.norm1_w; e1.ln_beta = L.norm1_b; e1.ln_eps = L.ln_eps;
    e1.ln_in = S.s1; e1.ln_stats = S.st1; e1.epilogue = 1;
    GM_TRY(lin(S.attn, d, n, d, L.out_proj_w, d, d, 0, L.out_proj_b, d, S.y, d, p, stream, &e1, L.p_out_proj));
    GM_TRY(lin(S.y, d, n, d, L.lin1_w, d, f, 0, L.lin1_b, f, S.u, f, p, stream, nullptr, L.p_lin1));
    geomae_linear_args e2{};
    e2.add_src = S.y; e2.ld_add = d; e2.ln_gamma = L.norm2_w; e2.ln_beta = L.norm2_b; e2.ln_eps = L.ln_eps;
    e2.ln_in = S.s2; e2.ln_stats = S.st2; e2.epilogue = 1; e2.a_gelu = 1;
    GM_TRY(lin(S.u, f, n, f, L.lin2_w, f, d, 0, L.lin2_b, d, S.z, d, p, stream, &e2, L.p_lin2));
    x = S.z;
  }
  return GEOMAE_OK;
}

namespace {

// Internal side streams / events, one set per device, created on first use.  Weight-gradient GEMMs run on a side
// stream so they overlap the latency-bound dX chain; the two decoder stacks run on separate streams.  The set holds
// no per-call state: every executor call re-synchronises the side streams with the caller's stream on entry
// (hand_off) and hands back on exit, so any number of models / trainers / host threads can share it; the event ring
// index is atomic, and an event is consumed (cudaStreamWaitEvent) right after it is recorded, before its slot can
// come round again.
struct Lanes {
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[64];
  std::atomic<unsigned> next_ev{0};
  bool ready = false;
  int init() {
    if (ready) return GEOMAE_OK;
    for (auto& st : side) GM_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& e : ev) GM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ready = true;
    return GEOMAE_OK;
  }
  cudaEvent_t event() { return ev[next_ev.fetch_add(1) % 64]; }
};
constexpr int MAX_DEVICES = 64;
Lanes g_lanes_of[MAX_DEVICES];
std::mutex g_lanes_mutex;

// the lanes of the CURRENT device (the one the caller's stream lives on); nullptr + error text on failure
Lanes* lanes() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) {
    gm_set_error("sra_stack: cannot identify the current device");
    return nullptr;
  }
  Lanes& l = g_lanes_of[dev];
  if (!l.ready) {
    std::lock_guard<std::mutex> lock(g_lanes_mutex);
    if (l.init() != GEOMAE_OK) return nullptr;
  }
  return &l;
}

// signal on `from`, wait on `to`
int hand_off(cudaStream_t from, cudaStream_t to) {
  Lanes* L_ = lanes();
  if (!L_) return GEOMAE_ERR_CUDA;
  cudaEvent_t e = L_->event();
  GM_CUDA(cudaEventRecord(e, from));
  GM_CUDA(cudaStreamWaitEvent(to, e, 0));
  return GEOMAE_OK;
}

int64_t scratch_floats(const geomae_sra_ctx* c) { return (int64_t)c->n_tokens * (9 * c->d_model + c->ffn + c->n_heads); }

// backward of one stack: dX chain on `main`, weight gradients on `side`; scratch holds TWO sets (alternating per
// layer) so the side stream may lag one layer behind.
int stack_backward_on(cudaStream_t main, cudaStream_t side, const geomae_sra_ctx* c, int32_t n_layers,
                      const geomae_sra_layer* layers, const geomae_sra_saved* saved, const float* x_in,
                      const float* d_out, float* d_in, float* scratch) {
  const int n = (int)c->n_tokens, d = c->d_model, f = c->ffn, p = c->precision;
  const int b16 = p == 1 ? 1 : 0;     // see geomae_sra_stack_forward
  const int64_t set = scratch_floats(c);
  cudaEvent_t side_done[2] = {nullptr, nullptr};
  if (p == 1) {
    // ---- bf16 mode: per layer ONE chain kernel (in-proj backward of the layer above .. out-proj backward, sra_chain.cu),
    // the attention backward, and ONE TMA-fed weight-gradient kernel (sra_wgrad.cu) on the side stream; a last chain
    // launch turns layer 0's dqkv into the input gradient.  Scratch set (per layer parity), in floats per token:
    // ds1 128 | dd 8 | ds2_16 64 | du16 128 | ds1_16 64 | dattn16 64 | dqkv16 192
    struct Set { float *ds1, *dd; void *ds2_16, *du16, *ds1_16, *dattn16, *dqkv16; };
    auto carve = [&](int parity) {
      float* b = scratch + (int64_t)parity * set;
      Set s;
      s.ds1 = b; b += (int64_t)n * 128;
      s.dd = b; b += (int64_t)n * 8;
      s.ds2_16 = b; b += (int64_t)n * 64;
      s.du16 = b; b += (int64_t)n * 128;
      s.ds1_16 = b; b += (int64_t)n * 64;
      s.dattn16 = b; b += (int64_t)n * 64;
      s.dqkv16 = b;
      return s;
    };
    const double dd_ = d, ff = f;
    const double up_flops = 2.0 * n * 3.0 * dd_ * dd_, up_bytes = n * (2.0 * 3.0 * dd_ + 4.0 * dd_) + 2.0 * 3.0 * dd_ * dd_;
    for (int l = n_layers - 1; l >= 0; --l) {
      const geomae_sra_layer& L = layers[l];
      const geomae_sra_saved& S = saved[l];
      const geomae_sra_windows& w = c->shift[L.shift];
      const Set cur = carve(l & 1);
      const bool top = l == n_layers - 1;
      if (side_done[l & 1]) GM_CUDA(cudaStreamWaitEvent(main, side_done[l & 1], 0));     // this set's last readers are done
      geomae_chain_bwd_args a{};
      a.n_tokens = n; a.mode = top ? 2 : 3;
      if (top) a.dz_in = d_out;
      else {
        const Set above = carve((l + 1) & 1);
        a.dqkv16_up = above.dqkv16; a.ds1_up = above.ds1; a.p_in_proj_up = layers[l + 1].p_in_proj[0];
      }
      a.xh2_16 = S.s2; a.st2 = S.st2; a.xh1_16 = S.s1; a.st1 = S.st1; a.u16 = S.u; a.attn16 = S.attn;
      a.p_lin2 = L.p_lin2[0]; a.p_lin1 = L.p_lin1[0]; a.p_out_proj = L.p_out_proj[0];
      a.norm2_w = L.norm2_w; a.norm1_w = L.norm1_w;
      a.ds2_16 = cur.ds2_16; a.du16 = cur.du16; a.ds1_16 = cur.ds1_16; a.dattn16 = cur.dattn16; a.ds1 = cur.ds1; a.dd = cur.dd;
      a.g_norm2_w = L.g_norm2_w; a.g_norm2_b = L.g_norm2_b; a.g_norm1_w = L.g_norm1_w; a.g_norm1_b = L.g_norm1_b;
      {
        // algorithmic bytes: dz (or dqkv' + ds1'), bf16 xhat2, xhat1, u, attn in; bf16 ds2, du, ds1, dattn + fp32 ds1 + D out; weights
        const double bytes = (top ? 4.0 * n * dd_ : up_bytes) + n * (2.0 * 2.0 * dd_ + 16.0 + 2.0 * ff + 2.0 * dd_) +
                             n * (3.0 * 2.0 * dd_ + 2.0 * ff + 4.0 * dd_ + 4.0 * c->n_heads) + 2.0 * (dd_ * dd_ + 2.0 * dd_ * ff);
        Span span(0, 2.0 * n * (dd_ * dd_ + 2.0 * dd_ * ff) + (top ? 0.0 : up_flops), main, bytes);
        GM_TRY(geomae_sra_chain_bwd(&a, main));
      }
      {
        Span span(3, 0.0, main, n * (2.0 * 3.0 * dd_ + 2.0 * dd_ + 8.0 * c->n_heads + 2.0 * 3.0 * dd_));
        GM_TRY(geomae_sra_attention_tc_bwd(S.qkv, S.attn, S.lse, (const float*)cur.dattn16, n, c->n_heads, w.win_ptr, w.win_tok,
                                           w.tok_win, (float*)cur.dqkv16, cur.dd, 1 | 2 | 4, main));
      }
      GM_TRY(hand_off(main, side));
      geomae_wgrad_layer_args g{};
      g.n_tokens = n;
      g.ds2_16 = cur.ds2_16; g.g16 = S.g; g.du16 = cur.du16; g.xh1_16 = S.s1; g.ds1_16 = cur.ds1_16; g.attn16 = S.attn;
      g.dqkv16 = cur.dqkv16; g.pos16 = c->pos16[L.shift];
      g.norm1_w = L.norm1_w; g.norm1_b = L.norm1_b;
      if (l == 0) { g.xin16 = saved[0].xb; g.in_scale = g.in_shift = nullptr; }             // the stack input itself
      else { g.xin16 = saved[l - 1].s2; g.in_scale = layers[l - 1].norm2_w; g.in_shift = layers[l - 1].norm2_b; }   // x = LN2 of the layer below
      g.g_lin2_w = L.g_lin2_w; g.g_lin1_w = L.g_lin1_w; g.g_lin1_b = L.g_lin1_b; g.g_out_proj_w = L.g_out_proj_w;
      g.g_in_proj_w = L.g_in_proj_w; g.g_in_proj_b = L.g_in_proj_b; g.g_lin2_b = L.g_lin2_b; g.g_out_proj_b = L.g_out_proj_b;
      {
        Span span(1, 2.0 * n * (3.0 * dd_ * dd_ + dd_ * dd_ + 2.0 * dd_ * ff), side, n * 2.0 * (6.0 * dd_ + 2.0 * ff + 3.0 * dd_) + 4.0 * (4.0 * dd_ * dd_ + 2.0 * dd_ * ff));
        GM_TRY(geomae_sra_wgrad_layer(&g, side));
      }
      side_done[l & 1] = lanes()->event();
      GM_CUDA(cudaEventRecord(side_done[l & 1], side));
    }
    {
      const Set s0 = carve(0);
      geomae_chain_bwd_args a{};
      a.n_tokens = n; a.mode = 1;
      a.dqkv16_up = s0.dqkv16; a.ds1_up = s0.ds1; a.p_in_proj_up = layers[0].p_in_proj[0]; a.dx = d_in;
      Span span(0, up_flops, main, up_bytes + 4.0 * n * dd_);
      GM_TRY(geomae_sra_chain_bwd(&a, main));
    }
    GM_TRY(hand_off(side, main));   // join
    return GEOMAE_OK;
  }
  struct StableGuard { ~StableGuard() { gm_set_weights_stable(false); } } guard;
  gm_set_weights_stable(true);      // backward re-uses the images packed by the forward pass of this step
  for (int l = n_layers - 1; l >= 0; --l) {
    float* base = scratch + (int64_t)(l & 1) * set;
    // ds2 | du | dy | ds1 | da | dqkv | dx | dd
    float* ds2 = base;
    float* du = ds2 + (int64_t)n * d;
    float* dy = du + (int64_t)n * f;
    float* ds1 = dy + (int64_t)n * d;
    float* da = ds1 + (int64_t)n * d;
    float* dqkv = da + (int64_t)n * d;
    float* dxb = dqkv + (int64_t)n * 3 * d;
    float* dd = dxb + (int64_t)n * d;
    const geomae_sra_layer& L = layers[l];
    const geomae_sra_saved& S = saved[l];
    const geomae_sra_windows& w = c->shift[L.shift];
    const float* x = l == 0 ? x_in : saved[l - 1].z;
    float* dx = l == 0 ? d_in : dxb;
    // ---- dX chain of the layer, back to back on `main` (no stream operations in between, so consecutive kernels
    // overlap through programmatic dependent launch).  Both LayerNorm backwards are GEMM epilogues (epilogue 3): the
    // gradient reaching a LayerNorm output never touches memory.  ds2 of this layer was written by the layer above
    // (its dx GEMM, below) — only the top layer runs the stand-alone LayerNorm backward on d_out.
    if (l == n_layers - 1) {
      if (side_done[l & 1]) GM_CUDA(cudaStreamWaitEvent(main, side_done[l & 1], 0));
      Span span(4, 0.0, main, 4.0 * n * (3.0 * d + 2.0));
      GM_TRY(geomae_layernorm_bwd(d_out, S.s2, S.st2, L.norm2_w, n, d, ds2, L.g_norm2_w, L.g_norm2_b, L.g_lin2_b, main));
    }
    geomae_linear_args e{};
    e.gelu_u = S.u; e.ldu = f; e.epilogue = 2; e.out_bf16 = b16;
    GM_TRY(lin(ds2, d, n, d, L.lin2_w, f, d, 1, nullptr, f, du, f, p, main, &e, L.p_lin2));
    geomae_linear_args e1{};           // dy = du W1 + ds2, then LayerNorm-1 backward in the epilogue -> ds1
    e1.add_src = ds2; e1.ld_add = d; e1.epilogue = 3;
    e1.ln_gamma = L.norm1_w; e1.ln_in = S.s1; e1.ln_stats = S.st1;
    e1.ln_dgamma = L.g_norm1_w; e1.ln_dbeta = L.g_norm1_b; e1.ln_dcolsum = L.g_out_proj_b;
    e1.a_bf16 = b16;
    GM_TRY(lin(du, f, n, f, L.lin1_w, d, f, 1, nullptr, d, ds1, d, p, main, &e1, L.p_lin1));
    geomae_linear_args ed{};          // bf16 mode: D = dO . O per (token, head) leaves this GEMM's epilogue
    if (p == 1) { ed.dot_src = S.attn; ed.ld_dot = d; ed.dot_out = dd; ed.out_bf16 = 1; }
    GM_TRY(lin(ds1, d, n, d, L.out_proj_w, d, d, 1, nullptr, d, da, d, p, main, &ed, L.p_out_proj));
    {
      Span span(3, 0.0, main, 4.0 * n * (3.0 * d + d + 2.0 * c->n_heads + 3.0 * d));
      if (p == 1)
        GM_TRY(geomae_sra_attention_tc_bwd(S.qkv, S.attn, S.lse, da, n, c->n_heads, w.win_ptr, w.win_tok, w.tok_win, dqkv,
                                           dd, 1 | 2 | 4, main));
      else
        GM_TRY(geomae_sra_attention_bwd(S.qkv, S.attn, S.lse, da, n, c->n_heads, w.win_ptr, w.win_tok, w.tok_win, dqkv,
                                        dd, main));
    }
    geomae_linear_args e2{};
    e2.add_src = ds1; e2.ld_add = d; e2.a_bf16 = b16;
    if (l > 0) {                       // dx = dqkv W_in + ds1 is the dz of the layer below: its LayerNorm-2 backward here
      const geomae_sra_layer& Lb = layers[l - 1];
      const geomae_sra_saved& Sb = saved[l - 1];
      float* ds2_below = scratch + (int64_t)((l - 1) & 1) * set;      // slot 0 of the other scratch set
      if (side_done[(l - 1) & 1]) GM_CUDA(cudaStreamWaitEvent(main, side_done[(l - 1) & 1], 0));   // that set is free again
      e2.epilogue = 3;
      e2.ln_gamma = Lb.norm2_w; e2.ln_in = Sb.s2; e2.ln_stats = Sb.st2;
      e2.ln_dgamma = Lb.g_norm2_w; e2.ln_dbeta = Lb.g_norm2_b; e2.ln_dcolsum = Lb.g_lin2_b;
      GM_TRY(lin(dqkv, 3 * d, n, 3 * d, L.in_proj_w, d, 3 * d, 1, nullptr, d, ds2_below, d, p, main, &e2, L.p_in_proj));
    } else {
      GM_TRY(lin(dqkv, 3 * d, n, 3 * d, L.in_proj_w, d, 3 * d, 1, nullptr, d, dx, d, p, main, &e2, L.p_in_proj));
    }
    // ---- the four weight gradients of the layer on `side`, one hand-off per layer; they overlap the next
    // layer's dX chain (bias gradients of linear2 / out_proj come from the LayerNorm backward above)
    GM_TRY(hand_off(main, side));
    GM_TRY(wgrad(ds2, d, S.u, f, n, L.g_lin2_w, f, nullptr, d, f, p, side, nullptr, nullptr, 0, 1));
    GM_TRY(wgrad(du, f, S.y, d, n, L.g_lin1_w, d, L.g_lin1_b, f, d, p, side, nullptr, nullptr, 0, 0, b16));
    GM_TRY(wgrad(ds1, d, S.attn, d, n, L.g_out_proj_w, d, nullptr, d, d, p, side));
    GM_TRY(wgrad(dqkv, 3 * d, x, d, n, L.g_in_proj_w, d, L.g_in_proj_b, 3 * d, d, p, side, c->pos_table, w.tok_cell, 2, 0, b16));
    side_done[l & 1] = lanes()->event();
    GM_CUDA(cudaEventRecord(side_done[l & 1], side));
  }
  GM_TRY(hand_off(side, main));   // join
  return GEOMAE_OK;
}

}  // namespace

extern "C" int64_t geomae_sra_scratch_floats(const geomae_sra_ctx* c, int32_t n_stacks) {
  return c ? 2 * n_stacks * scratch_floats(c) : 0;
}

extern "C" int geomae_sra_stack_backward(const geomae_sra_ctx* c, int32_t n_layers, const geomae_sra_layer* layers,
                                         const geomae_sra_saved* saved, const float* x_in, const float* d_out,
                                         float* d_in, float* scratch, void* stream) {
  GM_REQUIRE(c && layers && saved && d_out && d_in && scratch, "sra_stack_backward: null argument");
  if (c->n_tokens == 0) return GEOMAE_OK;
  Lanes* ln = lanes();
  if (!ln) return GEOMAE_ERR_CUDA;
  cudaStream_t main = (cudaStream_t)stream;
  GM_TRY(hand_off(main, ln->side[0]));     // the side stream must see everything queued before this call
  return stack_backward_on(main, ln->side[0], c, n_layers, layers, saved, x_in, d_out, d_in, scratch);
}

// Two stacks that read the same input (the centroid and density decoders, backbones/…top_only.py:269-277)
// executed concurrently on two streams.
extern "C" int geomae_sra_stack2_forward(const geomae_sra_ctx* c, int32_t n_layers, const geomae_sra_layer* layers_a,
                                         const geomae_sra_saved* saved_a, const geomae_sra_layer* layers_b,
                                         const geomae_sra_saved* saved_b, const float* x_in, void* stream) {
  Lanes* ln = lanes();
  if (!ln) return GEOMAE_ERR_CUDA;
  cudaStream_t main = (cudaStream_t)stream, other = ln->side[1];
  GM_TRY(hand_off(main, other));
  GM_TRY(geomae_sra_stack_forward(c, n_layers, layers_a, saved_a, x_in, main));
  GM_TRY(geomae_sra_stack_forward(c, n_layers, layers_b, saved_b, x_in, other));
  return hand_off(other, main);
}

extern "C" int geomae_sra_stack2_backward(const geomae_sra_ctx* c, int32_t n_layers, const geomae_sra_layer* layers_a,
                                          const geomae_sra_saved* saved_a, const geomae_sra_layer* layers_b,
                                          const geomae_sra_saved* saved_b, const float* x_in, const float* d_out_a,
                                          const float* d_out_b, float* d_in_a, float* d_in_b, float* scratch,
                                          void* stream) {
  GM_REQUIRE(c && layers_a && layers_b && saved_a && saved_b && d_out_a && d_out_b && d_in_a && d_in_b && scratch,
             "sra_stack2_backward: null argument");
  if (c->n_tokens == 0) return GEOMAE_OK;
  Lanes* ln = lanes();
  if (!ln) return GEOMAE_ERR_CUDA;
  cudaStream_t main = (cudaStream_t)stream;
  for (int i = 0; i < 3; ++i) GM_TRY(hand_off(main, ln->side[i]));
  GM_TRY(stack_backward_on(main, ln->side[0], c, n_layers, layers_a, saved_a, x_in, d_out_a, d_in_a, scratch));
  GM_TRY(stack_backward_on(ln->side[1], ln->side[2], c, n_layers, layers_b, saved_b, x_in, d_out_b, d_in_b,
                           scratch + 2 * scratch_floats(c)));
  return hand_off(ln->side[1], main);
}

// ---- per-family timing of the stack kernels (bench.py roofline leg)
extern "C" int geomae_profile_enable(int32_t on) {
  g_prof.on = on != 0;
  g_prof.n = 0;
  return GEOMAE_OK;
}

// Synchronises the device; ms[f], launches[f], flops[f] for the 5 families (see Prof).
extern "C" int geomae_profile_read(double* ms, int64_t* launches, double* flops, double* bytes) {
  GM_REQUIRE(ms && launches && flops, "profile_read: null argument");
  GM_CUDA(cudaDeviceSynchronize());
  for (int f = 0; f < Prof::FAMILIES; ++f) { ms[f] = 0.0; launches[f] = 0; flops[f] = 0.0; if (bytes) bytes[f] = 0.0; }
  for (int i = 0; i < g_prof.n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.start[i], g_prof.stop[i]) != cudaSuccess) continue;
    ms[g_prof.family[i]] += t;
    launches[g_prof.family[i]] += 1;
    flops[g_prof.family[i]] += g_prof.flops[i];
    if (bytes) bytes[g_prof.family[i]] += g_prof.bytes[i];
  }
  g_prof.n = 0;
  return GEOMAE_OK;
}
