// gw_train.inl -- training step of GraphWeatherForecaster on the exact-fp32 CUDA-core kernels: a forward that keeps what the
// backward needs, and the backward itself (every caller of the reference trains: /root/reference/train/run.py:508-543 calls
// loss.backward() on forecast.py:215-247).  Included by gw_api.cu (it works on gw_plan's graphs and weight views).
//
// The forward is the factored graph of gw_api.cu written out with every intermediate kept (h = hidden activations, z = the value
// entering LayerNorm); the backward walks it in reverse with five primitives (gw_simt.cu):
//     data gradient   dX = (dY . W) (.) relu-mask        the forward row-op kernel with the transposed weight and a mask
//     weight gradient dW += dY^T . A,  db += colsum(dY)   gw_wgrad_kernel (A assembled from a row source, like the forward)
//     LayerNorm       gw_ln_bwd_kernel                     dz, dgamma, dbeta
//     gathers         backward of x[src] / x[dst] / per-target sums: per-source / per-target segment sums and row gathers
//     broadcasts      tensors shared by the batch (encoded edge attributes, h3 node rows): gw_batch_reduce_kernel
// Gradients of the factored layer 1,  h1 = relu(e W1e^T + P_s[src] + P_d[dst] + b1),  P = x [W1s ; W1d]^T :
//     dP_s = per-source sum of dh1, dP_d = per-target sum of dh1;  dW1s += dP_s^T x, dW1d += dP_d^T x, dW1e += dh1^T e;
//     dx += dP_s W1s + dP_d W1d, de += dh1 W1e   -- two thirds of the K = 768 weight gradient are formed per NODE, not per edge.
// Scope: LayerNorm MLPs, node / edge / hidden dims <= 256, constraint_type "none" (the reference's default training setup).

namespace gw {

struct MlpTape {
  std::vector<float*> h;  // hidden activations h[l] [R, out_l], l < L
  float* z = nullptr;     // [R, out_L] value entering LayerNorm (null: MLP without norm)
  int rows = 0, batch = 0;
};

struct TrainState {
  cudaStream_t st = nullptr;
  std::vector<void*> allocs;  // stream-ordered allocations of the current step
  DevBuf<float> wT, gbuf;     // transposed weights / gradients, laid out like gw_plan::wbuf
  DevBuf<int32_t> lat_perm_src, lat_ptr_src, dec_perm_src, dec_ptr_src, iota;
  DevBuf<unsigned char> sort_ws;
  bool graphs_ready = false;
  int batch = 0;
  bool have_tape = false;
  // saved tensors of the current step
  const float* features = nullptr;
  float *xg = nullptr, *xm0 = nullptr, *e_enc = nullptr, *agg_m = nullptr, *e_lat = nullptr, *e_dec = nullptr, *agg_g = nullptr, *xg2 = nullptr;
  std::vector<float*> x, e, agg;  // x[k] k = 0..nb, e[k] k = 1..nb (e[0] = e_lat broadcast), agg[k]
  MlpTape t_enc_node_g, t_enc_node_h, t_enc_edge_enc, t_enc_edge, t_enc_mnode, t_lat_enc, t_dec_edge_enc, t_dec_edge, t_dec_node, t_dec_out;
  std::vector<MlpTape> t_pe, t_pn;
};

static float* talloc(TrainState* t, size_t floats) {
  void* q = nullptr;
  if (cudaMallocAsync(&q, std::max<size_t>(floats, 1) * sizeof(float), t->st) != cudaSuccess) return nullptr;
  t->allocs.push_back(q);
  return static_cast<float*>(q);
}
static void tfree_all(TrainState* t) {
  for (void* q : t->allocs) cudaFreeAsync(q, t->st);
  t->allocs.clear();
  t->have_tape = false;
}
#define GW_TALLOC(var, floats)                                          \
  float* var = talloc(T, (floats));                                     \
  GW_CHECK(var != nullptr, "training step: out of device memory")

static float* grad_of(gw_plan* p, TrainState* T, const float* w) { return T->gbuf.p + (w - p->wbuf.p); }
static const float* wT_of(gw_plan* p, TrainState* T, const float* w) { return T->wT.p + (w - p->wbuf.p); }

// MLP forward keeping h and z.  `first` describes Linear 0 (A sources / addends / weight slice / bias may be customised by the
// caller: factored layer 1); the result is out = residual + LN(...) (LN if the MLP has a norm).
static int mlp_fwd(gw_plan* p, TrainState* T, const Mlp& m, GemmOp first, const RowSrc& residual, float* out, int ldo, MlpTape* tape) {
  const int rows = first.rows_per_sample, batch = first.batch;
  const size_t R = (size_t)rows * batch;
  tape->rows = rows, tape->batch = batch;
  tape->h.assign(m.L, nullptr);
  for (int l = 0; l < m.L; ++l) {
    tape->h[l] = talloc(T, R * m.out[l]);
    GW_CHECK(tape->h[l] != nullptr, "training step: out of device memory");
  }
  const bool norm = m.ln_g != nullptr;
  if (norm) {
    tape->z = talloc(T, R * m.out[m.L]);
    GW_CHECK(tape->z != nullptr, "training step: out of device memory");
  }
  for (int l = 0; l <= m.L; ++l) {
    GemmOp op;
    if (l == 0) {
      op = first;
    } else {
      op.rows_per_sample = rows, op.batch = batch;
      op.a[0] = src_stream(tape->h[l - 1], m.in[l], m.in[l], rows);
      op.W = m.W[l], op.K = m.in[l], op.ldw = m.in[l], op.bias = m.b[l];
    }
    op.N = m.out[l];
    if (l < m.L) {
      op.relu = 1, op.out = tape->h[l], op.ldo = m.out[l];
    } else {
      op.relu = 0;
      if (norm) op.ln_gamma = m.ln_g, op.ln_beta = m.ln_b, op.save_pre = tape->z;
      op.residual = residual, op.out = out, op.ldo = ldo;
      GW_CHECK(!norm || ldo == m.out[l], "training MLP: LayerNorm output must be dense");
    }
    GW_TRY(run_op(p, op, T->st));
  }
  return 0;
}

// dX[R, N_in] = (dY[R, N_out] . W[N_out, N_in slice]) (.) (mask > 0) + add      W given as a view (pointer into wbuf, ld = ldw, col offset folded in)
static int dgrad(gw_plan* p, TrainState* T, const float* dY, int ldy, int n_out, int rows, int batch, const float* W_view, int w_rows, int w_cols_total,
                 int col0, int n_in, const float* mask, int ld_mask, const float* add, int ld_add, float* dX, int ldx) {
  // transposed weight: WT[k, n] for the whole matrix [w_cols_total, w_rows]; the slice starts at row col0
  const float* WT = wT_of(p, T, W_view) + (size_t)col0 * w_rows;
  GemmOp op;
  op.rows_per_sample = rows, op.batch = batch;
  op.a[0] = src_stream(dY, ldy, n_out, rows);
  op.W = WT, op.K = n_out, op.N = n_in, op.ldw = w_rows;
  if (add) op.add[0] = src_stream(add, ld_add, n_in, rows);
  if (mask) op.mask = src_stream(mask, ld_mask, n_in, rows);
  op.out = dX, op.ldo = ldx;
  (void)w_cols_total;
  return run_op(p, op, T->st);
}

// backward of an MLP from the gradient of its output down to the gradient at Linear 0's output (after the ReLU mask): dh0 [R, out_0]
static int mlp_bwd(gw_plan* p, TrainState* T, const Mlp& m, const MlpTape& tape, const float* dOut, int ld_dout, float** dh0_out) {
  const int rows = tape.rows, batch = tape.batch;
  const size_t R = (size_t)rows * batch;
  const float* cur = dOut;
  int ldc = ld_dout;
  if (tape.z) {
    GW_TALLOC(dz, R * m.out[m.L]);
    GW_CUDA(launch_ln_bwd(dOut, ld_dout, tape.z, m.out[m.L], m.out[m.L], m.ln_g, (long long)R, dz, m.out[m.L], grad_of(p, T, m.ln_g),
                          grad_of(p, T, m.ln_b), T->st));
    cur = dz, ldc = m.out[m.L];
  }
  for (int l = m.L; l >= 1; --l) {
    GW_CUDA(launch_wgrad(cur, ldc, m.out[l], src_stream(tape.h[l - 1], m.in[l], m.in[l], rows), m.in[l], rows, batch, grad_of(p, T, m.W[l]), m.in[l],
                         grad_of(p, T, m.b[l]), T->st));
    GW_TALLOC(dh, R * m.in[l]);
    GW_TRY(dgrad(p, T, cur, ldc, m.out[l], rows, batch, m.W[l], m.out[l], m.in[l], 0, m.in[l], tape.h[l - 1], m.in[l], nullptr, 0, dh, m.in[l]));
    cur = dh, ldc = m.in[l];
  }
  *dh0_out = const_cast<float*>(cur);
  return 0;
}

static int train_prepare(gw_plan* p, TrainState* T, int batch, cudaStream_t st) {
  const gw_dims& d = p->d;
  GW_CHECK(d.precision == GW_PREC_FP32_SIMT, "training runs on the exact-fp32 plan (precision fp32_simt)");
  GW_CHECK(p->w_enc && p->w_proc && p->w_dec && p->have_enc && p->have_lat && p->have_dec, "training needs the full forecaster (graphs + weights)");
  GW_CHECK(d.node_dim <= 256 && d.edge_dim <= 256 && d.hidden_node <= 256 && d.hidden_edge <= 256 && d.hidden_dec <= 256 && d.out_dim <= 256,
           "training kernels cover dims <= 256");
  T->st = st;
  if (T->wT.n != p->wbuf.n) GW_TRY(T->wT.alloc(p->wbuf.n));
  if (T->gbuf.n != p->wbuf.n) GW_TRY(T->gbuf.alloc(p->wbuf.n));
  for (const auto& kv : p->params) {  // transposed copies of every matrix (weights change every optimiser step)
    const int64_t r = kv.second.second.first, c = kv.second.second.second;
    if (c > 1) GW_CUDA(launch_transpose(kv.second.first, (int)r, (int)c, T->wT.p + (kv.second.first - p->wbuf.p), st));
  }
  if (!T->graphs_ready) {  // keep the stream-ordered pool's memory between steps (the default returns it to the driver at every sync)
    cudaMemPool_t pool = nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (!T->graphs_ready) {  // edges grouped by SOURCE (the x[src] gathers become per-source sums going backward)
    const int El = d.n_lat_edges, Ed = d.n_dec_edges, H = d.n_mesh;
    const size_t ws = std::max(sort_csr_workspace_bytes(El), sort_csr_workspace_bytes(Ed));
    GW_TRY(T->sort_ws.alloc(ws));
    GW_TRY(T->lat_perm_src.alloc(El) | T->lat_ptr_src.alloc(H + 1) | T->dec_perm_src.alloc(Ed) | T->dec_ptr_src.alloc(H + 1));
    GW_CUDA(launch_sort_csr(p->lat_src.p, El, H, T->lat_perm_src.p, T->lat_ptr_src.p, T->sort_ws.p, T->sort_ws.n, st));
    GW_CUDA(launch_sort_csr(p->dec_src.p, Ed, H, T->dec_perm_src.p, T->dec_ptr_src.p, T->sort_ws.p, T->sort_ws.n, st));
    T->graphs_ready = true;
  }
  T->batch = batch;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// forward (keeps activations)
// ---------------------------------------------------------------------------------------------------------------------------
static int train_forward(gw_plan* p, TrainState* T, const float* features, float* out, int B, cudaStream_t st) {
  const gw_dims& d = p->d;
  const int Dn = d.node_dim, De = d.edge_dim, He = d.hidden_edge, N = p->n_in_cur, H = d.n_mesh, El = d.n_lat_edges, Ed = d.n_dec_edges,
            No = d.n_out, nb = d.num_blocks;
  tfree_all(T);
  GW_TRY(train_prepare(p, T, B, st));
  p->cur_tag = TAG_CONST;
  RowSrc none;
  T->features = features;
  // ---- encoder ------------------------------------------------------------------------------------------------------------
  const Mlp& mne = p->enc_node;
  GW_TALLOC(xg, (size_t)B * N * Dn);
  GW_TRY(mlp_fwd(p, T, mne, first_op(N, B, src_stream(features, d.in_dim, d.in_dim, N), none, mne.W[0], mne.in[0], mne.in[0], mne.b[0]), none, xg, Dn,
                 &T->t_enc_node_g));
  GW_TALLOC(xm0, (size_t)H * Dn);
  GW_TRY(mlp_fwd(p, T, mne, first_op(H, 1, src_stream(p->h3_nodes, d.in_dim, d.in_dim, H), none, mne.W[0], mne.in[0], mne.in[0], mne.b[0]), none, xm0, Dn,
                 &T->t_enc_node_h));
  const Mlp& mee = p->enc_edge_enc;
  GW_TALLOC(e_enc, (size_t)N * De);
  GW_TRY(mlp_fwd(p, T, mee, first_op(N, 1, src_stream(p->enc_attr.p, d.enc_edge_attr_dim, d.enc_edge_attr_dim, N), none, mee.W[0], mee.in[0], mee.in[0],
                                     mee.b[0]), none, e_enc, De, &T->t_enc_edge_enc));
  // encoder block, edge MLP: h1 = relu(xg W1s^T + (xm0 W1d^T)[mesh] + e_enc W1e^T + b1)
  const Mlp& meb = p->enc_blk_edge;
  GW_TALLOC(Pm, (size_t)H * He);
  GW_TALLOC(Pe, (size_t)N * He);
  {
    GemmOp t;
    t.rows_per_sample = H, t.batch = 1, t.a[0] = src_stream(xm0, Dn, Dn, H), t.W = meb.W[0] + Dn, t.K = Dn, t.ldw = meb.in[0], t.N = He, t.out = Pm, t.ldo = He;
    GW_TRY(run_op(p, t, st));
    GemmOp c;
    c.rows_per_sample = N, c.batch = 1, c.a[0] = src_stream(e_enc, De, De, N), c.W = meb.W[0] + 2 * Dn, c.K = De, c.ldw = meb.in[0], c.N = He, c.out = Pe, c.ldo = He;
    GW_TRY(run_op(p, c, st));
  }
  GW_TALLOC(ep_enc, (size_t)B * N * De);
  {
    GemmOp fo = first_op(N, B, src_stream(xg, Dn, Dn, N), none, meb.W[0], Dn, meb.in[0], meb.b[0]);
    fo.add[0] = src_bgather(Pm, He, He, p->enc_mesh.p);
    fo.add[1] = src_bcast(Pe, He, He);
    GW_TRY(mlp_fwd(p, T, meb, fo, src_bcast(e_enc, De, De), ep_enc, De, &T->t_enc_edge));
  }
  GW_TALLOC(agg_m, (size_t)B * H * De);
  GW_CUDA(launch_segsum(ep_enc, De, De, p->enc_ptr.p, p->enc_perm.p, N, H, B, agg_m, De, st));
  const Mlp& mnb = p->enc_blk_node;
  T->x.assign(nb + 1, nullptr), T->e.assign(nb + 1, nullptr), T->agg.assign(nb, nullptr);
  T->t_pe.assign(nb, MlpTape()), T->t_pn.assign(nb, MlpTape());
  GW_TALLOC(x0, (size_t)B * H * Dn);
  GW_TRY(mlp_fwd(p, T, mnb, first_op(H, B, src_bcast(xm0, Dn, Dn), src_stream(agg_m, De, De, H), mnb.W[0], mnb.in[0], mnb.in[0], mnb.b[0]),
                 src_bcast(xm0, Dn, Dn), x0, Dn, &T->t_enc_mnode));
  T->x[0] = x0;
  const Mlp& mle = p->enc_lat_edge_enc;
  GW_TALLOC(e_lat, (size_t)El * De);
  GW_TRY(mlp_fwd(p, T, mle, first_op(El, 1, src_stream(p->lat_attr.p, 2, 2, El), none, mle.W[0], mle.in[0], mle.in[0], mle.b[0]), none, e_lat, De,
                 &T->t_lat_enc));
  T->xg = xg, T->xm0 = xm0, T->e_enc = e_enc, T->agg_m = agg_m, T->e_lat = e_lat;
  // ---- processor ----------------------------------------------------------------------------------------------------------
  GW_TALLOC(P, (size_t)B * H * 2 * He);
  for (int k = 0; k < nb; ++k) {
    const Mlp& me = p->proc_edge[k];
    const Mlp& mn = p->proc_node[k];
    for (int h = 0; h < 2; ++h) {
      GemmOp t;
      t.rows_per_sample = H, t.batch = B, t.a[0] = src_stream(T->x[k], Dn, Dn, H), t.W = me.W[0] + h * Dn, t.K = Dn, t.ldw = me.in[0], t.N = He;
      t.out = P + h * He, t.ldo = 2 * He;
      GW_TRY(run_op(p, t, st));
    }
    const RowSrc e_src = k == 0 ? src_bcast(e_lat, De, De) : src_stream(T->e[k], De, De, El);
    GW_TALLOC(en, (size_t)B * El * De);
    {
      GemmOp fo = first_op(El, B, e_src, none, me.W[0] + 2 * Dn, De, me.in[0], me.b[0]);
      fo.add[0] = src_gather(P, 2 * He, He, p->lat_src.p, H, 0);
      fo.add[1] = src_gather(P, 2 * He, He, p->lat_dst.p, H, He);
      GW_TRY(mlp_fwd(p, T, me, fo, e_src, en, De, &T->t_pe[k]));
    }
    T->e[k + 1] = en;
    GW_TALLOC(ag, (size_t)B * H * De);
    GW_CUDA(launch_segsum(en, De, De, p->lat_ptr.p, nullptr, El, H, B, ag, De, st));
    T->agg[k] = ag;
    GW_TALLOC(xn, (size_t)B * H * Dn);
    GW_TRY(mlp_fwd(p, T, mn, first_op(H, B, src_stream(T->x[k], Dn, Dn, H), src_stream(ag, De, De, H), mn.W[0], mn.in[0], mn.in[0], mn.b[0]),
                   src_stream(T->x[k], Dn, Dn, H), xn, Dn, &T->t_pn[k]));
    T->x[k + 1] = xn;
  }
  // ---- decoder ------------------------------------------------------------------------------------------------------------
  const Mlp& mde = p->dec_edge_enc;
  GW_TALLOC(e_dec, (size_t)Ed * De);
  GW_TRY(mlp_fwd(p, T, mde, first_op(Ed, 1, src_stream(p->dec_attr.p, 2, 2, Ed), none, mde.W[0], mde.in[0], mde.in[0], mde.b[0]), none, e_dec, De,
                 &T->t_dec_edge_enc));
  const Mlp& mdb = p->dec_blk_edge;
  GW_TALLOC(Pd, (size_t)B * H * He);
  {
    GemmOp t;
    t.rows_per_sample = H, t.batch = B, t.a[0] = src_stream(T->x[nb], Dn, Dn, H), t.W = mdb.W[0], t.K = Dn, t.ldw = mdb.in[0], t.N = He, t.out = Pd, t.ldo = He;
    GW_TRY(run_op(p, t, st));
  }
  GW_TALLOC(ep_dec, (size_t)B * Ed * De);
  {
    // h1 = relu(e_dec W1e^T + b1 + Pd[src])   (the lat/lon end of every decoder edge is a zero row: its W1d term vanishes)
    GemmOp fo = first_op(Ed, B, src_bcast(e_dec, De, De), none, mdb.W[0] + 2 * Dn, De, mdb.in[0], mdb.b[0]);
    fo.add[0] = src_gather(Pd, He, He, p->dec_src.p, H, 0);
    GW_TRY(mlp_fwd(p, T, mdb, fo, src_bcast(e_dec, De, De), ep_dec, De, &T->t_dec_edge));
  }
  GW_TALLOC(agg_g, (size_t)B * No * De);
  GW_CUDA(launch_segsum(ep_dec, De, De, p->dec_ptr.p, nullptr, Ed, No, B, agg_g, De, st));
  const Mlp& mdn = p->dec_blk_node;
  GW_TALLOC(xg2, (size_t)B * No * Dn);
  GW_TRY(mlp_fwd(p, T, mdn, first_op(No, B, src_stream(agg_g, De, De, No), none, mdn.W[0] + Dn, De, mdn.in[0], mdn.b[0]), none, xg2, Dn, &T->t_dec_node));
  const Mlp& mdo = p->dec_node_dec;
  RowSrc res;
  if (d.residual_dim > 0) res = src_stream(features, d.in_dim, d.out_dim, No);
  GW_TRY(mlp_fwd(p, T, mdo, first_op(No, B, src_stream(xg2, Dn, Dn, No), none, mdo.W[0], mdo.in[0], mdo.in[0], mdo.b[0]), res, out, d.out_dim, &T->t_dec_out));
  T->e_dec = e_dec, T->agg_g = agg_g, T->xg2 = xg2;
  T->have_tape = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------------
static int train_backward(gw_plan* p, TrainState* T, const float* dOut, float* dFeatures, cudaStream_t st) {
  const gw_dims& d = p->d;
  const int Dn = d.node_dim, De = d.edge_dim, He = d.hidden_edge, N = p->n_in_cur, H = d.n_mesh, El = d.n_lat_edges, Ed = d.n_dec_edges,
            No = d.n_out, nb = d.num_blocks, B = T->batch;
  GW_CHECK(T->have_tape, "gw_train_backward needs the activations of a preceding gw_train_forward");
  T->st = st;
  p->cur_tag = TAG_CONST;
  GW_CUDA(cudaMemsetAsync(T->gbuf.p, 0, T->gbuf.bytes(), st));
  if (dFeatures) GW_CUDA(cudaMemsetAsync(dFeatures, 0, (size_t)B * N * d.in_dim * sizeof(float), st));
  float* dh = nullptr;
  // ---- node_decoder (+ residual: d features[:, :out] = dOut) ------------------------------------------------------------------
  const Mlp& mdo = p->dec_node_dec;
  GW_TRY(mlp_bwd(p, T, mdo, T->t_dec_out, dOut, d.out_dim, &dh));
  GW_CUDA(launch_wgrad(dh, mdo.out[0], mdo.out[0], src_stream(T->xg2, Dn, Dn, No), Dn, No, B, grad_of(p, T, mdo.W[0]), mdo.in[0], grad_of(p, T, mdo.b[0]), st));
  GW_TALLOC(d_xg2, (size_t)B * No * Dn);
  GW_TRY(dgrad(p, T, dh, mdo.out[0], mdo.out[0], No, B, mdo.W[0], mdo.out[0], mdo.in[0], 0, Dn, nullptr, 0, nullptr, 0, d_xg2, Dn));
  // ---- decoder node MLP: xg2 = LN(MLP(agg_g with W1[:, Dn:]))   (the x half of the concat is identically zero) ---------------
  const Mlp& mdn = p->dec_blk_node;
  GW_TRY(mlp_bwd(p, T, mdn, T->t_dec_node, d_xg2, Dn, &dh));
  GW_CUDA(launch_wgrad(dh, mdn.out[0], mdn.out[0], src_stream(T->agg_g, De, De, No), De, No, B, grad_of(p, T, mdn.W[0]) + Dn, mdn.in[0],
                       grad_of(p, T, mdn.b[0]), st));
  GW_TALLOC(d_agg_g, (size_t)B * No * De);
  GW_TRY(dgrad(p, T, dh, mdn.out[0], mdn.out[0], No, B, mdn.W[0], mdn.out[0], mdn.in[0], Dn, De, nullptr, 0, nullptr, 0, d_agg_g, De));
  // ---- decoder edge MLP: e' = LN(...) + e_dec, agg_g = per-point sum of e' ------------------------------------------------------
  const Mlp& mdb = p->dec_blk_edge;
  GW_TALLOC(d_ep, (size_t)B * Ed * De);
  GW_CUDA(launch_gather_rows(d_agg_g, De, No, p->dec_dst.p, Ed, De, B, d_ep, De, false, st));
  GW_TALLOC(d_e_dec, (size_t)Ed * De);
  GW_CUDA(launch_batch_reduce(d_ep, De, Ed, De, B, d_e_dec, De, false, st));  // residual: e_dec is shared by the batch
  GW_TRY(mlp_bwd(p, T, mdb, T->t_dec_edge, d_ep, De, &dh));                   // dh = dh1 [B*Ed, He]
  {
    // layer 0: h1 = relu(e_dec W1e^T + b1 + Pd[src]); e_dec is shared by the batch, so its terms use the batch-reduced gradient
    GW_TALLOC(dPe, (size_t)Ed * He);
    GW_CUDA(launch_batch_reduce(dh, He, Ed, He, B, dPe, He, false, st));
    GW_CUDA(launch_wgrad(dPe, He, He, src_stream(T->e_dec, De, De, Ed), De, Ed, 1, grad_of(p, T, mdb.W[0]) + 2 * Dn, mdb.in[0], grad_of(p, T, mdb.b[0]), st));
    GW_TRY(dgrad(p, T, dPe, He, He, Ed, 1, mdb.W[0], mdb.out[0], mdb.in[0], 2 * Dn, De, nullptr, 0, d_e_dec, De, d_e_dec, De));
    GW_TALLOC(dPd, (size_t)B * H * He);
    GW_CUDA(launch_segsum(dh, He, He, T->dec_ptr_src.p, T->dec_perm_src.p, Ed, H, B, dPd, He, st));
    GW_CUDA(launch_wgrad(dPd, He, He, src_stream(T->x[nb], Dn, Dn, H), Dn, H, B, grad_of(p, T, mdb.W[0]), mdb.in[0], nullptr, st));
    GW_TALLOC(dx_last, (size_t)B * H * Dn);
    GW_TRY(dgrad(p, T, dPd, He, He, H, B, mdb.W[0], mdb.out[0], mdb.in[0], 0, Dn, nullptr, 0, nullptr, 0, dx_last, Dn));
    dh = dx_last;
  }
  float* dx = dh;  // gradient of x[nb]
  // decoder.edge_encoder
  const Mlp& mde = p->dec_edge_enc;
  {
    float* g0 = nullptr;
    GW_TRY(mlp_bwd(p, T, mde, T->t_dec_edge_enc, d_e_dec, De, &g0));
    GW_CUDA(launch_wgrad(g0, mde.out[0], mde.out[0], src_stream(p->dec_attr.p, 2, 2, Ed), 2, Ed, 1, grad_of(p, T, mde.W[0]), mde.in[0], grad_of(p, T, mde.b[0]), st));
  }
  // ---- processor blocks, last to first ------------------------------------------------------------------------------------------
  float* de = nullptr;  // gradient of e[k+1] (none flows into the last block's e')
  for (int k = nb - 1; k >= 0; --k) {
    const Mlp& me = p->proc_edge[k];
    const Mlp& mn = p->proc_node[k];
    // node MLP: x[k+1] = LN(MLP([x[k] ; agg[k]])) + x[k]
    GW_TRY(mlp_bwd(p, T, mn, T->t_pn[k], dx, Dn, &dh));
    GW_CUDA(launch_wgrad(dh, mn.out[0], mn.out[0], src_stream(T->x[k], Dn, Dn, H), Dn, H, B, grad_of(p, T, mn.W[0]), mn.in[0], grad_of(p, T, mn.b[0]), st));
    GW_CUDA(launch_wgrad(dh, mn.out[0], mn.out[0], src_stream(T->agg[k], De, De, H), De, H, B, grad_of(p, T, mn.W[0]) + Dn, mn.in[0], nullptr, st));
    GW_TALLOC(dxk, (size_t)B * H * Dn);
    GW_TRY(dgrad(p, T, dh, mn.out[0], mn.out[0], H, B, mn.W[0], mn.out[0], mn.in[0], 0, Dn, nullptr, 0, dx, Dn, dxk, Dn));  // + residual path
    GW_TALLOC(d_agg, (size_t)B * H * De);
    GW_TRY(dgrad(p, T, dh, mn.out[0], mn.out[0], H, B, mn.W[0], mn.out[0], mn.in[0], Dn, De, nullptr, 0, nullptr, 0, d_agg, De));
    // e[k+1] receives its target's aggregate gradient (+ what the next block sent back)
    GW_TALLOC(d_en, (size_t)B * El * De);
    if (de) {
      GW_CUDA(cudaMemcpyAsync(d_en, de, (size_t)B * El * De * sizeof(float), cudaMemcpyDeviceToDevice, st));
      GW_CUDA(launch_gather_rows(d_agg, De, H, p->lat_dst.p, El, De, B, d_en, De, true, st));
    } else {
      GW_CUDA(launch_gather_rows(d_agg, De, H, p->lat_dst.p, El, De, B, d_en, De, false, st));
    }
    // edge MLP: e[k+1] = LN(...) + e[k];  h1 = relu(e[k] W1e^T + P_s[src] + P_d[dst] + b1)
    GW_TRY(mlp_bwd(p, T, me, T->t_pe[k], d_en, De, &dh));
    const RowSrc e_src = k == 0 ? src_bcast(T->e_lat, De, De) : src_stream(T->e[k], De, De, El);
    GW_CUDA(launch_wgrad(dh, He, He, e_src, De, El, B, grad_of(p, T, me.W[0]) + 2 * Dn, me.in[0], grad_of(p, T, me.b[0]), st));
    GW_TALLOC(d_ek, (size_t)B * El * De);
    GW_TRY(dgrad(p, T, dh, He, He, El, B, me.W[0], me.out[0], me.in[0], 2 * Dn, De, nullptr, 0, d_en, De, d_ek, De));  // + residual path
    GW_TALLOC(dPs, (size_t)B * H * He);
    GW_TALLOC(dPt, (size_t)B * H * He);
    GW_CUDA(launch_segsum(dh, He, He, T->lat_ptr_src.p, T->lat_perm_src.p, El, H, B, dPs, He, st));
    GW_CUDA(launch_segsum(dh, He, He, p->lat_ptr.p, nullptr, El, H, B, dPt, He, st));
    GW_CUDA(launch_wgrad(dPs, He, He, src_stream(T->x[k], Dn, Dn, H), Dn, H, B, grad_of(p, T, me.W[0]), me.in[0], nullptr, st));
    GW_CUDA(launch_wgrad(dPt, He, He, src_stream(T->x[k], Dn, Dn, H), Dn, H, B, grad_of(p, T, me.W[0]) + Dn, me.in[0], nullptr, st));
    GW_TALLOC(dx1, (size_t)B * H * Dn);
    GW_TRY(dgrad(p, T, dPs, He, He, H, B, me.W[0], me.out[0], me.in[0], 0, Dn, nullptr, 0, dxk, Dn, dx1, Dn));
    GW_TALLOC(dx2, (size_t)B * H * Dn);
    GW_TRY(dgrad(p, T, dPt, He, He, H, B, me.W[0], me.out[0], me.in[0], Dn, Dn, nullptr, 0, dx1, Dn, dx2, Dn));
    dx = dx2, de = d_ek;
  }
  // e[0] = e_lat broadcast: reduce over the batch, back through latent_edge_encoder
  {
    const Mlp& mle = p->enc_lat_edge_enc;
    GW_TALLOC(d_elat, (size_t)El * De);
    GW_CUDA(launch_batch_reduce(de, De, El, De, B, d_elat, De, false, st));
    float* g0 = nullptr;
    GW_TRY(mlp_bwd(p, T, mle, T->t_lat_enc, d_elat, De, &g0));
    GW_CUDA(launch_wgrad(g0, mle.out[0], mle.out[0], src_stream(p->lat_attr.p, 2, 2, El), 2, El, 1, grad_of(p, T, mle.W[0]), mle.in[0], grad_of(p, T, mle.b[0]), st));
  }
  // ---- encoder block, node MLP (mesh rows): x[0] = LN(MLP([xm0 ; agg_m])) + xm0 ------------------------------------------------
  const Mlp& mnb = p->enc_blk_node;
  GW_TALLOC(d_xm0, (size_t)H * Dn);
  GW_CUDA(launch_batch_reduce(dx, Dn, H, Dn, B, d_xm0, Dn, false, st));  // residual: xm0 is shared by the batch
  GW_TRY(mlp_bwd(p, T, mnb, T->t_enc_mnode, dx, Dn, &dh));
  GW_CUDA(launch_wgrad(dh, mnb.out[0], mnb.out[0], src_bcast(T->xm0, Dn, Dn), Dn, H, B, grad_of(p, T, mnb.W[0]), mnb.in[0], grad_of(p, T, mnb.b[0]), st));
  GW_CUDA(launch_wgrad(dh, mnb.out[0], mnb.out[0], src_stream(T->agg_m, De, De, H), De, H, B, grad_of(p, T, mnb.W[0]) + Dn, mnb.in[0], nullptr, st));
  {
    GW_TALLOC(t1, (size_t)B * H * Dn);
    GW_TRY(dgrad(p, T, dh, mnb.out[0], mnb.out[0], H, B, mnb.W[0], mnb.out[0], mnb.in[0], 0, Dn, nullptr, 0, nullptr, 0, t1, Dn));
    GW_CUDA(launch_batch_reduce(t1, Dn, H, Dn, B, d_xm0, Dn, true, st));
  }
  GW_TALLOC(d_agg_m, (size_t)B * H * De);
  GW_TRY(dgrad(p, T, dh, mnb.out[0], mnb.out[0], H, B, mnb.W[0], mnb.out[0], mnb.in[0], Dn, De, nullptr, 0, nullptr, 0, d_agg_m, De));
  // ---- encoder block, edge MLP (lat/lon rows): e' = LN(...) + e_enc;  h1 = relu(xg W1s^T + Pm[mesh] + Pe + b1) ------------------
  const Mlp& meb = p->enc_blk_edge;
  GW_TALLOC(d_epe, (size_t)B * N * De);
  GW_CUDA(launch_gather_rows(d_agg_m, De, H, p->enc_mesh.p, N, De, B, d_epe, De, false, st));
  GW_TALLOC(d_e_enc, (size_t)N * De);
  GW_CUDA(launch_batch_reduce(d_epe, De, N, De, B, d_e_enc, De, false, st));
  GW_TRY(mlp_bwd(p, T, meb, T->t_enc_edge, d_epe, De, &dh));
  GW_CUDA(launch_wgrad(dh, He, He, src_stream(T->xg, Dn, Dn, N), Dn, N, B, grad_of(p, T, meb.W[0]), meb.in[0], grad_of(p, T, meb.b[0]), st));
  GW_TALLOC(d_xg, (size_t)B * N * Dn);
  GW_TRY(dgrad(p, T, dh, He, He, N, B, meb.W[0], meb.out[0], meb.in[0], 0, Dn, nullptr, 0, nullptr, 0, d_xg, Dn));
  {
    GW_TALLOC(dPm_b, (size_t)B * H * He);
    GW_CUDA(launch_segsum(dh, He, He, p->enc_ptr.p, p->enc_perm.p, N, H, B, dPm_b, He, st));
    GW_TALLOC(dPm, (size_t)H * He);
    GW_CUDA(launch_batch_reduce(dPm_b, He, H, He, B, dPm, He, false, st));
    GW_CUDA(launch_wgrad(dPm, He, He, src_stream(T->xm0, Dn, Dn, H), Dn, H, 1, grad_of(p, T, meb.W[0]) + Dn, meb.in[0], nullptr, st));
    GW_TRY(dgrad(p, T, dPm, He, He, H, 1, meb.W[0], meb.out[0], meb.in[0], Dn, Dn, nullptr, 0, d_xm0, Dn, d_xm0, Dn));
    GW_TALLOC(dPe, (size_t)N * He);
    GW_CUDA(launch_batch_reduce(dh, He, N, He, B, dPe, He, false, st));
    GW_CUDA(launch_wgrad(dPe, He, He, src_stream(T->e_enc, De, De, N), De, N, 1, grad_of(p, T, meb.W[0]) + 2 * Dn, meb.in[0], nullptr, st));
    GW_TRY(dgrad(p, T, dPe, He, He, N, 1, meb.W[0], meb.out[0], meb.in[0], 2 * Dn, De, nullptr, 0, d_e_enc, De, d_e_enc, De));
  }
  // encoder.edge_encoder, node_encoder on the h3 rows and on the lat/lon rows
  {
    const Mlp& mee = p->enc_edge_enc;
    float* g0 = nullptr;
    GW_TRY(mlp_bwd(p, T, mee, T->t_enc_edge_enc, d_e_enc, De, &g0));
    GW_CUDA(launch_wgrad(g0, mee.out[0], mee.out[0], src_stream(p->enc_attr.p, d.enc_edge_attr_dim, d.enc_edge_attr_dim, N), d.enc_edge_attr_dim, N, 1,
                         grad_of(p, T, mee.W[0]), mee.in[0], grad_of(p, T, mee.b[0]), st));
    const Mlp& mne = p->enc_node;
    GW_TRY(mlp_bwd(p, T, mne, T->t_enc_node_h, d_xm0, Dn, &g0));
    GW_CUDA(launch_wgrad(g0, mne.out[0], mne.out[0], src_stream(p->h3_nodes, d.in_dim, d.in_dim, H), d.in_dim, H, 1, grad_of(p, T, mne.W[0]), mne.in[0],
                         grad_of(p, T, mne.b[0]), st));
    if (p->params.count("encoder.h3_nodes"))  // h3_nodes is a learned parameter of the forecaster (encoder.py:113)
      GW_TRY(dgrad(p, T, g0, mne.out[0], mne.out[0], H, 1, mne.W[0], mne.out[0], mne.in[0], 0, d.in_dim, nullptr, 0, nullptr, 0,
                   grad_of(p, T, p->h3_nodes), d.in_dim));
    GW_TRY(mlp_bwd(p, T, mne, T->t_enc_node_g, d_xg, Dn, &g0));
    GW_CUDA(launch_wgrad(g0, mne.out[0], mne.out[0], src_stream(T->features, d.in_dim, d.in_dim, N), d.in_dim, N, B, grad_of(p, T, mne.W[0]), mne.in[0],
                         grad_of(p, T, mne.b[0]), st));
    if (dFeatures) {
      GW_TRY(dgrad(p, T, g0, mne.out[0], mne.out[0], N, B, mne.W[0], mne.out[0], mne.in[0], 0, d.in_dim, nullptr, 0, nullptr, 0, dFeatures, d.in_dim));
      if (d.residual_dim > 0)  // out = node_decoder(...) + features[..., :out]: the residual passes dOut straight through
        GW_CUDA(launch_strided_add(dOut, d.out_dim, dFeatures, d.in_dim, (long long)B * N, d.out_dim, st));
    }
  }
  (void)No;
  tfree_all(T);
  return 0;
}

}  // namespace gw
