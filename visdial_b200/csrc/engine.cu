// Engine orchestration: replaces Model:forwardBackward / retrieveBatch (/root/reference/model.lua:249-430)
// and the nn graphs of encoders/*.lua + decoders/*.lua for the four configured encoders.
#include "engine.h"
#include <string.h>
#include <math.h>

namespace vd {

// ------------------------------------------------------------------------------------------------
// config + parameter layout
// ------------------------------------------------------------------------------------------------
Cfg parse_cfg(const vd_params* p) {
  VD_REQUIRE(p && p->encoder && p->decoder, VD_E_BADARG, "params / encoder / decoder is null");
  Cfg c;
  c.encoder = p->encoder; c.decoder = p->decoder;
  struct { const char* name; int enc; } known[] = {
      {"lf-ques", ENC_LF_QUES}, {"lf-ques-im-hist", ENC_LF_QIH}, {"hrea-ques-im-hist", ENC_HREA}, {"mn-att-ques-im-hist", ENC_MN_ATT},
      {"lf-ques-im", ENC_LF_QI}, {"lf-ques-hist", ENC_LF_QH}, {"hre-ques-hist", ENC_HRE_QH}, {"hre-ques-im-hist", ENC_HRE_QIH},
      {"mn-ques-hist", ENC_MN_QH}, {"mn-ques-im-hist", ENC_MN_QIH}, {"lf-att-ques-im-hist", ENC_LF_ATT}};
  c.enc = -1;
  for (auto& k : known)
    if (c.encoder == k.name) c.enc = k.enc;
  VD_REQUIRE(c.enc >= 0, VD_E_BADARG, "unknown encoder (one of the eleven names of encoders/*.lua)");
  if (c.decoder == "disc") c.dec = DEC_DISC;
  else if (c.decoder == "gen") c.dec = DEC_GEN;
  else VD_REQUIRE(false, VD_E_BADARG, "unknown decoder (disc | gen)");
  c.V = p->vocabSize; c.E = p->embedSize; c.H = p->rnnHiddenSize; c.L = p->numLayers; c.F = p->imgFeatureSize;
  c.S = p->imgSpatialSize; c.IE = p->imgEmbedSize; c.Cm = p->commonEmbeddingSize; c.hops = p->numAttentionLayers;
  c.R = p->maxQuesCount; c.K = p->numOptions; c.dropout = p->dropout; c.gpuid = p->gpuid;
  // opts.lua:55-59,62
  c.useHist = c.encoder.find("hist") != std::string::npos;
  c.useIm = c.encoder.find("im") != std::string::npos;
  c.att = c.encoder.find("att") != std::string::npos;
  c.fam_lf = c.encoder.compare(0, 3, "lf-") == 0;
  c.fam_hre = c.encoder.compare(0, 3, "hre") == 0;
  c.fam_mn = c.encoder.compare(0, 3, "mn-") == 0;
  c.hre_att = c.enc == ENC_HREA;
  c.san = c.att;
  c.img_in_q = c.fam_hre && c.useIm;
  c.img_drop = c.enc == ENC_HREA;
  c.mn_qi = c.enc == ENC_MN_QIH;
  c.embdrop = c.fam_mn || c.enc == ENC_LF_ATT;
  c.rnn_layers = (c.fam_lf && !c.att) || c.fam_hre;
  if (c.enc == ENC_LF_ATT) c.hops = 1;           // hard-wired upstream: lf-att-ques-im-hist.lua:49 does not read params.numAttentionLayers
  VD_REQUIRE(c.V > 2 && c.E > 0 && c.H > 0, VD_E_BADARG, "vocabSize / embedSize / rnnHiddenSize must be positive");
  VD_REQUIRE(c.L == 2, VD_E_BADARG, "numLayers must be 2");
  VD_REQUIRE(c.E % 4 == 0 && c.H % 32 == 0 && c.IE % 4 == 0 && c.Cm % 4 == 0 && c.F % 4 == 0, VD_E_BADARG,
             "embedSize/imgEmbedSize/commonEmbeddingSize/imgFeatureSize must be multiples of 4, rnnHiddenSize of 32");
  VD_REQUIRE(c.R >= 1 && c.R <= 32 && c.K >= 1 && c.K <= 1024 && c.hops >= 1, VD_E_BADARG, "maxQuesCount/numOptions/numAttentionLayers out of range");
  VD_REQUIRE(c.dropout >= 0.f && c.dropout < 1.f, VD_E_BADARG, "dropout must be in [0,1)");
  return c;
}

int Layout::find(const std::string& name) const {
  for (size_t i = 0; i < segs.size(); ++i)
    if (segs[i].name == name) return (int)i;
  return -1;
}

static void add_seg(Layout& l, const std::string& name, int64_t rows, int64_t cols, int init, int64_t fan_in) {
  Seg s; s.name = name; s.rows = rows; s.cols = cols; s.init = init; s.fan_in = fan_in; s.off = l.total;
  l.segs.push_back(s);
  l.total += (rows * cols + 31) / 32 * 32;       // every segment starts 128-byte aligned
}
static void add_lstm(Layout& l, const std::string& n, int D, int H) {
  add_seg(l, n + ".weight", D + H, 4 * H, VD_INIT_LSTM_W, D + H);
  add_seg(l, n + ".bias", 1, 4 * H, VD_INIT_LSTM_B, D + H);
}
static void add_linear(Layout& l, const std::string& n, int out, int in) {
  add_seg(l, n + ".weight", out, in, VD_INIT_LINEAR_W, in);
  add_seg(l, n + ".bias", 1, out, VD_INIT_LINEAR_B, in);
}

Layout build_layout(const Cfg& c) {
  Layout l;
  add_seg(l, "wordEmbed.weight", c.V + 1, c.E, VD_INIT_EMBED, 0);
  auto add_san = [&]() {
    add_linear(l, "san.img", c.H, c.F);
    for (int h = 1; h <= c.hops; ++h) {
      std::string p = "san.hop" + std::to_string(h) + ".";
      add_linear(l, p + "img_common", c.Cm, c.H);
      add_linear(l, p + "ques_common", c.Cm, c.H);
      add_linear(l, p + "score", 1, c.Cm);
    }
    add_linear(l, "san.out", c.H, c.H);
  };
  if (c.fam_lf && !c.san) {
    // lf-ques / lf-ques-im / lf-ques-hist / lf-ques-im-hist: fusion over [q | img | h]
    add_lstm(l, "ques.lstm1", c.E, c.H); add_lstm(l, "ques.lstm2", c.H, c.H);
    if (c.useHist) { add_lstm(l, "hist.lstm1", c.E, c.H); add_lstm(l, "hist.lstm2", c.H, c.H); }
    add_linear(l, "fusion", c.H, c.H + (c.useIm ? c.F : 0) + (c.useHist ? c.H : 0));
  } else if (c.fam_hre) {
    if (c.img_in_q) add_linear(l, "img.embed", c.IE, c.F);
    add_lstm(l, "hist.lstm1", c.E, c.H); add_lstm(l, "hist.lstm2", c.H, c.H);
    add_lstm(l, "ques.lstm1", c.E + (c.img_in_q ? c.IE : 0), c.H); add_lstm(l, "ques.lstm2", c.H, c.H);
    if (c.hre_att) { add_linear(l, "att.q", 1, c.H); add_linear(l, "att.h", 1, c.H); }
    add_lstm(l, "dialog.lstm", 2 * c.H, c.H);
  } else {
    // mn-* and lf-att-ques-im-hist: history then question LSTMs, then the attention blocks
    add_lstm(l, "hist.lstm1", c.E, c.H); add_lstm(l, "hist.lstm2", c.H, c.H);
    add_lstm(l, "ques.lstm1", c.E, c.H); add_lstm(l, "ques.lstm2", c.H, c.H);
    if (c.fam_lf) add_linear(l, "fusion", c.H, 2 * c.H);                   // lf-att: tanh(Linear([q | h]))
    if (c.mn_qi) add_linear(l, "mn.qi", c.H, c.H + c.F);                   // mn-ques-im-hist: tanh(Linear([q | fc7]))
    if (c.fam_mn) { add_linear(l, "mn.fact", c.H, c.H); add_linear(l, "mn.query", c.H, c.H); }
    if (c.san) add_san();
  }
  if (c.dec == DEC_DISC) {
    add_lstm(l, "opt.lstm", c.E, c.H);
  } else {
    add_lstm(l, "dec.lstm1", c.E, c.H); add_lstm(l, "dec.lstm2", c.H, c.H);
    add_linear(l, "dec.out", c.V, c.H);
  }
  return l;
}

// ------------------------------------------------------------------------------------------------
// memory
// ------------------------------------------------------------------------------------------------
void* Arena::alloc(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  while (cur < chunks.size()) {
    Chunk& c = chunks[cur];
    if (c.used + bytes <= c.cap) { void* p = c.p + c.used; c.used += bytes; return p; }
    ++cur;
  }
  Chunk c; c.cap = std::max<size_t>(bytes, (size_t)256 << 20); c.used = bytes;
  VD_CUDA_CHECK(cudaMalloc((void**)&c.p, c.cap));
  chunks.push_back(c);
  cur = chunks.size() - 1;
  return c.p;
}
void Arena::reset() { for (auto& c : chunks) c.used = 0; cur = 0; }
void Arena::release() { for (auto& c : chunks) cudaFree(c.p); chunks.clear(); cur = 0; }

void* GrowBuf::ensure(size_t bytes) {
  if (bytes > cap) {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    VD_CUDA_CHECK(cudaMalloc(&p, bytes));
    cap = bytes;
  }
  return p;
}
void GrowBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

// ------------------------------------------------------------------------------------------------
Engine::Engine(const vd_params* p) {
  cfg = parse_cfg(p);
  lay = build_layout(cfg);
  nparams = lay.total;
  int ndev = 0;
  VD_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  VD_REQUIRE(cfg.gpuid >= 0 && cfg.gpuid < ndev, VD_E_BADARG, "gpuid out of range");
  VD_CUDA_CHECK(cudaSetDevice(cfg.gpuid));
  cudaDeviceProp prop;
  VD_CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg.gpuid));
  VD_REQUIRE(prop.major == 10, VD_E_CUDA, "visdial_b200 is built for sm_100a (B200) only");
  cx.sm_count = prop.multiProcessorCount;
  // the encoder's chains of small dependent kernels get the highest priority, the option LSTM's SM-filling launches
  // the lowest: a freed SM goes to the latency-bound chain first
  int prio_lo = 0, prio_hi = 0;
  VD_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  VD_CUDA_CHECK(cudaStreamCreateWithPriority(&cx.stream, cudaStreamNonBlocking, prio_hi));
  main_stream = cx.stream;
  VD_CUDA_CHECK(cudaStreamCreateWithPriority(&side_stream, cudaStreamNonBlocking, prio_hi));
  VD_CUDA_CHECK(cudaStreamCreateWithPriority(&main2_stream, cudaStreamNonBlocking, prio_hi));
  VD_CUDA_CHECK(cudaStreamCreateWithPriority(&side2_stream, cudaStreamNonBlocking, prio_hi));
  VD_CUDA_CHECK(cudaStreamCreateWithPriority(&main3_stream, cudaStreamNonBlocking, prio_hi));
  VD_CUDA_CHECK(cudaStreamCreateWithPriority(&side3_stream, cudaStreamNonBlocking, prio_hi));
  VD_CUDA_CHECK(cudaStreamCreateWithPriority(&opt_stream, cudaStreamNonBlocking, prio_lo));
  VD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  VD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  VD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_opt_fork, cudaEventDisableTiming));
  VD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_opt_done, cudaEventDisableTiming));
  if (const char* s = getenv("VD_OPT_OVERLAP")) opt_overlap = atoi(s) != 0;
  if (const char* s = getenv("VD_OPT_RESERVE_SMS")) opt_reserve_sms = std::max(0, std::min(atoi(s), cx.sm_count - 16)) & ~1;
  size_t bytes = (size_t)nparams * sizeof(float);
  VD_CUDA_CHECK(cudaMalloc((void**)&W, bytes));
  VD_CUDA_CHECK(cudaMalloc((void**)&dW, bytes));
  VD_CUDA_CHECK(cudaMalloc((void**)&m, bytes));
  VD_CUDA_CHECK(cudaMalloc((void**)&v, bytes));
  VD_CUDA_CHECK(cudaMalloc((void**)&Wt, bytes));
  VD_CUDA_CHECK(cudaMemsetAsync(W, 0, bytes, cx.stream));
  VD_CUDA_CHECK(cudaMemsetAsync(dW, 0, bytes, cx.stream));
  VD_CUDA_CHECK(cudaMemsetAsync(m, 0, bytes, cx.stream));
  VD_CUDA_CHECK(cudaMemsetAsync(v, 0, bytes, cx.stream));
  VD_CUDA_CHECK(cudaMemsetAsync(Wt, 0, bytes, cx.stream));
  VD_CUDA_CHECK(cudaMalloc((void**)&scalars_dev, 64 * sizeof(float)));
  // table of 2-D weight segments that get a transposed shadow copy
  std::vector<int64_t> tab;
  for (auto& s : lay.segs) {
    if (s.init == VD_INIT_LSTM_W || s.init == VD_INIT_LINEAR_W) {
      tab.push_back(s.off); tab.push_back(s.rows); tab.push_back(s.cols);
      max2d = std::max(max2d, s.rows * s.cols);
    }
  }
  nseg2d = (int)tab.size() / 3;
  VD_CUDA_CHECK(cudaMalloc((void**)&segtab_dev, tab.size() * sizeof(int64_t)));
  VD_CUDA_CHECK(cudaMemcpyAsync(segtab_dev, tab.data(), tab.size() * sizeof(int64_t), cudaMemcpyHostToDevice, cx.stream));
  VD_CUDA_CHECK(cudaStreamSynchronize(cx.stream));
  VD_CUDA_CHECK(cudaEventCreate(&t0));
  VD_CUDA_CHECK(cudaEventCreate(&t1));
}

Engine::~Engine() {
  cudaSetDevice(cfg.gpuid);
  cudaStreamSynchronize(cx.stream);
  if (opt_stream) cudaStreamSynchronize(opt_stream);
  arena.release();
  for (auto& g : stage) g.release();
  if (copy_stream) {
    cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream);
    cudaEventDestroy(ev_img_ready); cudaEventDestroy(ev_copy_fork);
    for (auto ev : ev_img_free) cudaEventDestroy(ev);
  }
  for (auto& g : stage_img) g.release();
  cudaFree(W); cudaFree(dW); cudaFree(m); cudaFree(v); cudaFree(Wt); cudaFree(scalars_dev); cudaFree(segtab_dev);
  if (flush_buf) cudaFree(flush_buf);
  cx.collect();
  for (auto e : cx.free_events) cudaEventDestroy(e);
  if (t0) cudaEventDestroy(t0);
  if (t1) cudaEventDestroy(t1);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (side_stream) { cudaStreamSynchronize(side_stream); cudaStreamDestroy(side_stream); }
  if (main2_stream) { cudaStreamSynchronize(main2_stream); cudaStreamDestroy(main2_stream); }
  if (side2_stream) { cudaStreamSynchronize(side2_stream); cudaStreamDestroy(side2_stream); }
  if (main3_stream) { cudaStreamSynchronize(main3_stream); cudaStreamDestroy(main3_stream); }
  if (side3_stream) { cudaStreamSynchronize(side3_stream); cudaStreamDestroy(side3_stream); }
  if (opt_stream) { cudaStreamSynchronize(opt_stream); cudaStreamDestroy(opt_stream); }
  if (ev_opt_fork) cudaEventDestroy(ev_opt_fork);
  if (ev_opt_done) cudaEventDestroy(ev_opt_done);
  for (auto e : ev_pool) cudaEventDestroy(e);
  cudaStreamDestroy(main_stream);
}

int Engine::seg(const char* name) const {
  int i = lay.find(name);
  VD_REQUIRE(i >= 0, VD_E_STATE, name);
  return i;
}

DropCfg Engine::dropcfg(float p) const {
  DropCfg d;
  d.seed_lo = (uint32_t)drop_seed; d.seed_hi = (uint32_t)(drop_seed >> 32); d.iter = (uint32_t)drop_iter;
  if (training == 1 && p > 0.f) {
    double t = (double)p * 4294967296.0;
    d.thresh = (uint32_t)std::min(t, 4294967295.0);
    d.scale = 1.f / (1.f - p);
  } else { d.thresh = 0; d.scale = 1.f; }
  return d;
}

void Engine::gemm_tn(int M, int N, int K, const float* A, int64_t lda, const int32_t* gather, const float* B, int64_t ldb,
                     float* C, int64_t ldc, float beta, const float* bias, int act) {
  LaunchCtx::Scope sc(&cx, "gemm", 2.0 * M * N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N));
  if (tcmode() && gemm_tn_tc(cx, M, N, K, A, lda, gather, B, ldb, C, ldc, beta, bias, act)) return;
  gemm_tn_simt(cx, M, N, K, A, lda, gather, B, ldb, C, ldc, beta, bias, act);
}
void Engine::gemm_atb(int M, int N, int64_t K, const float* A, int64_t lda, const int32_t* gather, const float* B,
                      int64_t ldb, float* C, int64_t ldc) {
  LaunchCtx::Scope sc(&cx, "gemm_wgrad", 2.0 * M * N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N));
  if (tcmode() && gemm_atb_tc(cx, M, N, K, A, lda, gather, B, ldb, C, ldc)) return;
  gemm_atb_simt(cx, M, N, K, A, lda, gather, B, ldb, C, ldc);
}

void Engine::linear_fwd(int wseg, const float* x, int64_t rows, float* y, int act) {
  const Seg& s = lay.segs[wseg];
  gemm_tn((int)rows, (int)s.rows, (int)s.cols, x, s.cols, nullptr, Wp(wseg), s.cols, y, s.rows, 0.f, Wp(wseg + 1), act);
}
void Engine::linear_bwd(int wseg, const float* x, const float* dy, int64_t rows, float* dx, float beta_dx) {
  const Seg& s = lay.segs[wseg];
  int out = (int)s.rows, in = (int)s.cols;
  gemm_atb(out, in, rows, dy, out, nullptr, x, in, dWp(wseg), in);
  colsum_add(cx, dWp(wseg + 1), dy, rows, out, out);
  if (dx) gemm_tn((int)rows, in, out, dy, out, nullptr, Wtp(wseg), out, dx, in, beta_dx, nullptr, 0);
}

void Engine::refresh_shadows() {
  // LookupTableMaskZero.updateOutput zeroes the pad row at every forward [upstream rnn]
  VD_CUDA_CHECK(cudaMemsetAsync(W, 0, (size_t)cfg.E * sizeof(float), cx.stream));
  transpose_segments(cx, W, Wt, segtab_dev, nseg2d, max2d);
}

void Engine::stage_batch(const vd_batch* b) {
  VD_REQUIRE(b != nullptr, VD_E_BADARG, "batch is null");
  VD_REQUIRE(b->B > 0, VD_E_SHAPE, "batch B must be > 0");
  db = DevBatch();
  db.B = b->B; db.N = (int64_t)b->B * cfg.R;
  db.Tq = b->Tq; db.Th = b->Th; db.Ta = b->Ta; db.To = b->To;
  VD_REQUIRE(b->ques_fwd && b->Tq > 0, VD_E_SHAPE, "ques_fwd / Tq missing");
  if (cfg.useHist) VD_REQUIRE(b->hist && b->Th > 0, VD_E_SHAPE, "encoder uses history: hist / Th missing");
  if (cfg.useIm) VD_REQUIRE(b->img_feat, VD_E_SHAPE, "encoder uses the image: img_feat missing");
  int64_t img_elems = cfg.att ? (int64_t)b->B * cfg.S * cfg.S * cfg.F : (int64_t)b->B * cfg.F;
  struct Item { const void* src; size_t bytes; const void** dst; } items[9] = {
      {b->ques_fwd, (size_t)db.N * b->Tq * 4, (const void**)&db.ques},
      {cfg.useHist ? b->hist : nullptr, (size_t)db.N * b->Th * 4, (const void**)&db.hist},
      {cfg.useIm ? b->img_feat : nullptr, (size_t)img_elems * 4, (const void**)&db.img},
      {b->options, (size_t)db.N * cfg.K * b->To * 4, (const void**)&db.options},
      {b->answer_ind, (size_t)db.N * 4, (const void**)&db.answer_ind},
      {b->answer_in, (size_t)db.N * b->Ta * 4, (const void**)&db.answer_in},
      {b->answer_out, (size_t)db.N * b->Ta * 4, (const void**)&db.answer_out},
      {b->option_in, (size_t)db.N * cfg.K * b->To * 4, (const void**)&db.option_in},
      {b->option_out, (size_t)db.N * cfg.K * b->To * 4, (const void**)&db.option_out}};
  img_copy_pending = false;
  for (int i = 0; i < 9; ++i) {
    if (!items[i].src || items[i].bytes == 0) { *items[i].dst = nullptr; continue; }
    if (b->on_device) { *items[i].dst = items[i].src; continue; }
    if (i == 2 && items[i].bytes >= ((size_t)1 << 20)) {
      // image features: asynchronous copy on the copy stream into the staging buffer the previous step is not using
      if (!copy_stream) {
        VD_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        VD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_img_ready, cudaEventDisableTiming));
        VD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_copy_fork, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) VD_CUDA_CHECK(cudaEventCreateWithFlags(&ev_img_free[k], cudaEventDisableTiming));
      }
      img_slot ^= 1;
      if (items[i].bytes > stage_img[img_slot].cap) VD_CUDA_CHECK(cudaStreamSynchronize(copy_stream));   // growing frees the old block
      void* d = stage_img[img_slot].ensure(items[i].bytes);
      VD_CUDA_CHECK(cudaEventRecord(ev_copy_fork, cx.stream));                  // stream order behind whatever freed / allocated before
      VD_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_copy_fork, 0));
      if (img_free_recorded[img_slot]) VD_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_img_free[img_slot], 0));
      VD_CUDA_CHECK(cudaMemcpyAsync(d, items[i].src, items[i].bytes, cudaMemcpyHostToDevice, copy_stream));
      VD_CUDA_CHECK(cudaEventRecord(ev_img_ready, copy_stream));
      img_copy_pending = true;
      *items[i].dst = d;
      continue;
    }
    void* d = stage[i].ensure(items[i].bytes);
    VD_CUDA_CHECK(cudaMemcpyAsync(d, items[i].src, items[i].bytes, cudaMemcpyHostToDevice, cx.stream));
    *items[i].dst = d;
  }
}

void Engine::wait_img() {
  if (!img_copy_pending) return;
  VD_CUDA_CHECK(cudaStreamWaitEvent(cx.stream, ev_img_ready, 0));
  if (cx.stream != main_stream) VD_CUDA_CHECK(cudaStreamWaitEvent(main_stream, ev_img_ready, 0));   // later readers live on the main stream
  img_copy_pending = false;
}
void Engine::release_img() {
  if (!copy_stream || !db.img || db.img != stage_img[img_slot].p) return;
  VD_CUDA_CHECK(cudaEventRecord(ev_img_free[img_slot], cx.stream));
  img_free_recorded[img_slot] = true;
}

// ------------------------------------------------------------------------------------------------
// nn.SeqLSTM [upstream rnn], SURVEY.md Appendix C.  Forward: one batched x-projection (or per step in
// the non-saving mode) + per step {recurrent GEMM, pointwise}.
// ------------------------------------------------------------------------------------------------
void Engine::lstm_forward_begin(LstmRun& r, bool save) {
  const Seg& ws = lay.segs[r.wseg];
  VD_REQUIRE(ws.rows == r.D + r.H && ws.cols == 4 * r.H, VD_E_STATE, "lstm weight shape");
  const int H = r.H, D = r.D, G = 4 * r.H;
  const int64_t R = r.R;
  const float* WtS = Wtp(r.wseg);          // [4H, D+H]
  const float* bias = Wp(r.wseg + 1);
  const float* A = r.x ? r.x : Wp(0);
  const int64_t lda = r.x ? D : cfg.E;
  r.saved = save;
  // Tensor-core path: the x-projection of an embedding-gathered input becomes a (V+1, 4H) projection table
  // computed once per forward (E Wx^T: the 300-wide half of every step's contraction collapses into a
  // gather in the step epilogue); a dense input keeps the batched x-projection.  Each step is then ONE fused
  // kernel: recurrent tcgen05 GEMM + SeqLSTM pointwise epilogue.
  r.tc = tcmode() && H % 64 == 0;
  r.ptable = nullptr;
  // VD_MATH_F16: a many-row LSTM over embedding-gathered tokens (the option LSTM) keeps h, the activated gates, da and
  // the projection table in fp16 and runs kind::f16 contractions (lstm16.cu); c, the accumulators and h_T stay fp32
  r.f16 = math_mode == VD_MATH_F16 && r.gather && !r.x && !r.h0 && D == cfg.E && lstm16_shape_ok(R, H);
  if (r.f16) {
    float* pt = arena.get<float>((int64_t)(cfg.V + 1) * G);
    gemm_tn(cfg.V + 1, G, D, Wp(0), cfg.E, nullptr, WtS, D + H, pt, G, 0.f, nullptr, 0);     // bias stays fp32, added per step
    r.P16 = arena.get<__half>((int64_t)(cfg.V + 1) * G);
    cvt_f32_to_f16(cx, r.P16, G, pt, G, cfg.V + 1, G);
    r.Wh16 = arena.get<__half>((int64_t)G * H);
    cvt_f32_to_f16(cx, r.Wh16, H, WtS + D, D + H, G, H);                                    // [4H, H]: B operand of the forward step
    r.Whb16 = nullptr;
    if (save) {
      r.Whb16 = arena.get<__half>((int64_t)H * G);
      cvt_f32_to_f16(cx, r.Whb16, G, Wp(r.wseg) + (int64_t)D * G, G, H, G);                 // [H, 4H]: B operand of the backward step
    }
    const int64_t slots = save ? r.T : 2;
    r.h16 = arena.get<__half>(slots * R * H);
    r.c = arena.get<float>(slots * R * H);
    r.gates16 = save ? arena.get<__half>((int64_t)r.T * R * G) : nullptr;
    r.h32_last = arena.get<float>(R * H);
    r.h = nullptr; r.gates = nullptr;
    return;
  }
  // on the tensor-core path the bias is folded into the x-projection (table or GEMM epilogue), so the per-step
  // kernel reads one array less
  const float* xbias = r.tc ? bias : nullptr;
  if (r.tc && r.gather) {
    float* pt = arena.get<float>((int64_t)(cfg.V + 1) * G);
    gemm_tn(cfg.V + 1, G, D, Wp(0), cfg.E, nullptr, WtS, D + H, pt, G, 0.f, xbias, 0);
    r.ptable = pt;
  }
  if (save) {
    r.h = arena.get<float>((int64_t)r.T * R * H);
    r.c = arena.get<float>((int64_t)r.T * R * H);
    r.gates = arena.get<float>((int64_t)r.T * R * G);
    if (!r.ptable && !(r.tc && r.step_xproj))
      gemm_tn((int)((int64_t)r.T * R), G, D, A, lda, r.gather, WtS, D + H, r.gates, G, 0.f, xbias, 0);
  } else {
    r.h = arena.get<float>(2 * R * H);
    r.c = arena.get<float>(2 * R * H);
    r.gates = arena.get<float>(R * G);
  }
}

void Engine::lstm_forward_step(LstmRun& r, int t) {
  const int H = r.H, D = r.D, G = 4 * r.H;
  const int64_t R = r.R;
  const bool save = r.saved;
  const float* WtS = Wtp(r.wseg);
  const float* bias = Wp(r.wseg + 1);
  const float* A = r.x ? r.x : Wp(0);
  const int64_t lda = r.x ? D : cfg.E;
  const float* xbias = r.tc ? bias : nullptr;
  const int64_t slot = save ? t : (t & 1), pslot = save ? t - 1 : ((t - 1) & 1);
  float* g = save ? r.gates + (int64_t)t * R * G : r.gates;
  const float* hp = t > 0 ? r.h + pslot * R * H : r.h0;
  const float* cp = t > 0 ? r.c + pslot * R * H : r.c0;
  const bool per_step_x = !r.ptable && (!save || (r.tc && r.step_xproj));
  if (r.f16) {
    __half* g16 = save ? r.gates16 + (int64_t)t * R * G : nullptr;
    const int32_t* mk = r.mask ? r.mask + (int64_t)t * R : nullptr;
    const int32_t* tok = r.gather + (int64_t)t * R;
    float* h32 = t == r.T - 1 ? r.h32_last : nullptr;
    if (t == 0) {       // no recurrent term: a streaming kernel, accounted outside the roofline class
      LaunchCtx::Scope sc(&cx, "lstm_step_first", 0.0, R * (2.0 * G + 2.0 * G + 6.0 * H));
      lstm16_first_step(cx, R, H, r.P16, tok, bias, cp, mk, g16, r.c + slot * R * H, r.h16 + slot * R * H, h32);
    } else {
      // algorithmic HBM bytes: fp16 gates out, fp32 c in + out, fp16 h in + out (the fp16 table gather is L2-resident: not counted)
      LaunchCtx::Scope sc(&cx, "lstm_step", 2.0 * R * G * H, R * (2.0 * G + 4.0 * H + 4.0 * H + 2.0 * H + 2.0 * H));
      lstm16_step_fwd(cx, R, H, r.h16 + pslot * R * H, r.Wh16, r.P16, tok, bias, cp, mk, g16, r.c + slot * R * H,
                      r.h16 + slot * R * H, h32);
    }
    return;
  }
  // the big (option-LSTM) launches run alone on the GPU: they are the roofline kernel class; the 320-row encoder
  // steps overlap on 4 streams and are accounted separately
  // EXECUTED work only: a gathered x-projection is a table lookup, and a first step without initial state has no
  // recurrent contraction at all (it is a streaming kernel, kept out of the roofline class)
  const bool first_no_rec = r.tc && !hp;
  LaunchCtx::Scope sc(&cx, first_no_rec ? "lstm_step_first" : (R >= 4096 ? "lstm_step" : "lstm_step_small"),
                      first_no_rec ? 0.0 : 2.0 * R * G * (H + (per_step_x ? D : 0)), 4.0 * R * (G + 4.0 * H));
  if (r.tc) {
    const int32_t* mk = r.mask ? r.mask + (int64_t)t * R : nullptr;
    int has_x = 0;
    if (!r.ptable) {
      if (per_step_x && !r.xproj_external)
        gemm_tn((int)R, G, D, r.x + (int64_t)t * R * D, lda, nullptr, WtS, D + H, g, G, 0.f, xbias, 0);
      has_x = 1;
    }
    const int32_t* tok = r.ptable ? r.gather + (int64_t)t * R : nullptr;
    if (!hp) {          // t = 0 without h0: no recurrent term, a plain streaming kernel (x-projection already has the bias)
      lstm_first_step_fwd(cx, (save || has_x) ? g : nullptr, r.ptable, tok, nullptr, cp, mk, r.c + slot * R * H,
                          r.h + slot * R * H, R, H);
      return;
    }
    bool ok = lstm_step_fwd_tc(cx, R, H, hp, WtS + D, D + H, nullptr, (save || has_x) ? g : nullptr, has_x, r.ptable, tok, cp,
                               r.c + slot * R * H, r.h + slot * R * H, mk);
    VD_REQUIRE(ok, VD_E_STATE, "lstm_step_fwd_tc refused a shape the engine routed to it");
    return;
  }
  if (!save) {
    const float* At = r.x ? r.x + (int64_t)t * R * D : A;
    gemm_tn((int)R, G, D, At, lda, r.gather ? r.gather + (int64_t)t * R : nullptr, WtS, D + H, g, G, 0.f, nullptr, 0);
  }
  if (hp) gemm_tn((int)R, G, H, hp, H, nullptr, WtS + D, D + H, g, G, 1.f, nullptr, 0);
  lstm_pointwise_fwd(cx, g, bias, cp, r.mask ? r.mask + (int64_t)t * R : nullptr, r.c + slot * R * H, r.h + slot * R * H, R, H);
}

// the per-step x-projection of a `step_xproj` run, issued by the caller on the stream of its choice (cx.stream)
void Engine::lstm_forward_xproj(LstmRun& r, int t) {
  const int H = r.H, D = r.D, G = 4 * r.H;
  const int64_t R = r.R;
  VD_REQUIRE(r.tc && r.step_xproj && r.saved && r.x, VD_E_STATE, "lstm_forward_xproj: not a per-step projected run");
  gemm_tn((int)R, G, D, r.x + (int64_t)t * R * D, D, nullptr, Wtp(r.wseg), D + H, r.gates + (int64_t)t * R * G, G, 0.f,
          Wp(r.wseg + 1), 0);
}

void Engine::lstm_forward(LstmRun& r, bool save) {
  lstm_forward_begin(r, save);
  for (int t = 0; t < r.T; ++t) lstm_forward_step(r, t);
}

cudaEvent_t Engine::pool_event(size_t i) {
  while (ev_pool.size() <= i) {
    cudaEvent_t e;
    VD_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ev_pool.push_back(e);
  }
  return ev_pool[i];
}

// Wavefront over two stacked layers: layer-2 step t only needs layer-1 step t, so the two recurrences run one step
// apart on two streams instead of back to back (the per-step kernels of these 320-row LSTMs are latency-bound).
// Below this many rows the per-step contractions leave the tensor-core path (M < 64) and the two-stream wavefront only
// adds launches and event traffic.
static int64_t wavefront_min_rows() {
  static int64_t v = -1;
  if (v < 0) { const char* e = getenv("VD_WAVEFRONT_MIN_ROWS"); v = e ? atoll(e) : 64; }
  return v;
}

static bool three_streams_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VD_THREE_STREAMS"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

static bool enc_persist_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VD_ENC_PERSIST"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

void Engine::lstm_pair_forward(LstmRun& l1, LstmRun& l2, cudaStream_t sa, cudaStream_t sb, cudaStream_t sc) {
  // VD_MATH_F16: both layers and all time steps in ONE persistent launch (enc_lstm.cu) — weight slices stationary in
  // shared memory, steps chained through global flags — instead of 3 launches per time step on three streams
  if (math_mode == VD_MATH_F16 && enc_persist_enabled() && !l1.h0 && !l1.c0 && !l2.h0 && !l2.c0 && !l1.gather && l1.x &&
      l1.H == l2.H && l2.D == l1.H && l1.R == l2.R && l1.T == l2.T && l1.mask == l2.mask && enc_pair_shape_ok(l1.R, l1.H, cx.sm_count)) {
    const int H = l1.H, G = 4 * H, T = l1.T;
    const int64_t R = l1.R;
    cx.stream = sa;
    lstm_forward_begin(l1, true);               // h / c / gates of layer 1; gates1 <- x-projection + bias (batched GEMM)
    l2.x = l1.h;
    l2.step_xproj = true;                       // no batched x-projection for layer 2: the kernel's projection CTAs compute it per step
    lstm_forward_begin(l2, true);
    l1.h16 = arena.get<__half>((int64_t)T * R * H);
    l2.h16 = arena.get<__half>((int64_t)T * R * H);
    l2.x16 = l1.h16;
    __half* W1h = arena.get<__half>((int64_t)G * H);
    __half* W2c = arena.get<__half>((int64_t)G * 2 * H);
    cvt_f32_to_f16(cx, W1h, H, Wtp(l1.wseg) + l1.D, l1.D + H, G, H);                   // [4H, H]   recurrent block of layer 1
    cvt_f32_to_f16(cx, W2c, 2 * H, Wtp(l2.wseg), 2 * H, G, 2 * H);                     // [4H, 2H]  [Wx2 | Wh2] of layer 2
    int* flags = arena.get<int>(3 * (int64_t)cdiv(R, 128) * T);
    LaunchCtx::Scope sc2(&cx, "enc_pair_fwd", 2.0 * T * R * G * 3.0 * H, 4.0 * T * R * (2.0 * G + 2.0 * G + 6.0 * H));
    enc_pair_forward(cx, T, R, H, W1h, W2c, Wp(l2.wseg + 1), l1.mask, l1.gates, l1.c, l1.h, l1.h16, l2.gates, l2.c, l2.h, l2.h16, flags);
    return;
  }
  const bool pipelined = tcmode() && l2.H % 64 == 0 && sb != nullptr && sb != sa && l2.R >= wavefront_min_rows();
  const bool three = pipelined && sc != nullptr && sc != sa && sc != sb && three_streams_enabled();
  cx.stream = sa;
  if (!pipelined) {
    lstm_forward(l1, true);
    l2.x = l1.h;
    lstm_forward(l2, true);
    return;
  }
  lstm_forward_begin(l1, true);
  l2.x = l1.h;
  l2.step_xproj = true;
  cudaEvent_t e0 = pool_event(0);
  VD_CUDA_CHECK(cudaEventRecord(e0, sa));
  VD_CUDA_CHECK(cudaStreamWaitEvent(sb, e0, 0));
  cx.stream = sb;
  lstm_forward_begin(l2, true);
  l2.xproj_external = three;
  for (int t = 0; t < l1.T; ++t) {
    cx.stream = sa;
    lstm_forward_step(l1, t);
    cudaEvent_t e = pool_event(1 + t);
    VD_CUDA_CHECK(cudaEventRecord(e, sa));
    if (three) {
      // layer 2's x-projection of step t needs h1_t only: on its own stream it runs beside layer 2's step t-1, so
      // each of the three chains advances by ONE kernel per time step
      VD_CUDA_CHECK(cudaStreamWaitEvent(sc, e, 0));
      if (t == 0) VD_CUDA_CHECK(cudaStreamWaitEvent(sc, e0, 0));
      cx.stream = sc;
      lstm_forward_xproj(l2, t);
      cudaEvent_t ex = pool_event(1 + l1.T + t);
      VD_CUDA_CHECK(cudaEventRecord(ex, sc));
      VD_CUDA_CHECK(cudaStreamWaitEvent(sb, ex, 0));
    } else {
      VD_CUDA_CHECK(cudaStreamWaitEvent(sb, e, 0));
    }
    cx.stream = sb;
    lstm_forward_step(l2, t);
  }
  VD_CUDA_CHECK(cudaEventRecord(e0, sb));
  VD_CUDA_CHECK(cudaStreamWaitEvent(sa, e0, 0));
  cx.stream = sa;
}

void Engine::lstm_backward_begin(LstmRun& r, const float* dh_all, const float* dh_last, const float* dc_last) {
  VD_REQUIRE(r.saved, VD_E_STATE, "lstm_backward needs a forward run in training mode");
  const int H = r.H, G = 4 * r.H;
  const int64_t R = r.R, TR = (int64_t)r.T * R;
  r.dc_carry = arena.get<float>(R * H);
  if (!r.f16) {
    r.da = arena.get<float>(TR * G);
    r.dh_rec = arena.get<float>(R * H);
  }
  r.bw_dh_all = dh_all; r.bw_dh_last = dh_last; r.bw_dc_last = dc_last;
  if (r.f16) {
    VD_REQUIRE(dh_last && !dh_all && !dc_last, VD_E_STATE, "fp16 BPTT takes its gradient from the last step only");
    r.da16 = arena.get<__half>(TR * G);
    r.scale2 = arena.get<float>(4);
    pick_grad_scale(cx, dh_last, R * H, reinterpret_cast<uint32_t*>(r.scale2 + 2), r.scale2);
    return;
  }
  r.bw_tc = tcmode() && H % 128 == 0;
  if (r.bw_tc && dc_last)
    VD_CUDA_CHECK(cudaMemcpyAsync(r.dc_carry, dc_last, (size_t)R * H * sizeof(float), cudaMemcpyDeviceToDevice, cx.stream));
  else
    VD_CUDA_CHECK(cudaMemsetAsync(r.dc_carry, 0, (size_t)R * H * sizeof(float), cx.stream));
}

void Engine::lstm_backward_step(LstmRun& r, int t) {
  const int H = r.H, D = r.D, G = 4 * r.H;
  const int64_t R = r.R;
  const float* Ws = Wp(r.wseg);            // (D+H, 4H)
  const float* cp = t > 0 ? r.c + (int64_t)(t - 1) * R * H : r.c0;
  const int32_t* mk = r.mask ? r.mask + (int64_t)t * R : nullptr;
  float* da_t = r.da + (int64_t)t * R * G;
  const bool last = t == r.T - 1;
  const float* ext = r.bw_dh_all ? r.bw_dh_all + (int64_t)t * R * H : nullptr;
  if (r.f16) {
    const __half* g16 = r.gates16 + (int64_t)t * R * G;
    __half* da16_t = r.da16 + (int64_t)t * R * G;
    if (last) {
      LaunchCtx::Scope sc(&cx, "lstm_step_bwd_last", 0.0, R * (2.0 * G + 2.0 * G + 16.0 * H));
      lstm16_bwd_last(cx, R, H, g16, cp, r.c + (int64_t)t * R * H, r.bw_dh_last, r.scale2, mk, r.dc_carry, da16_t);
    } else {
      LaunchCtx::Scope sc(&cx, "lstm_step_bwd", 2.0 * R * G * H, R * (3.0 * 2.0 * G + 16.0 * H));
      lstm16_step_bwd(cx, R, H, r.da16 + (int64_t)(t + 1) * R * G, r.Whb16, g16, cp, r.c + (int64_t)t * R * H, r.dc_carry, mk, da16_t);
    }
    return;
  }
  LaunchCtx::Scope sc(&cx, (last && r.bw_tc) ? "lstm_step_bwd_last" : (R >= 4096 ? "lstm_step_bwd" : "lstm_step_bwd_small"), last ? 0.0 : 2.0 * R * G * H, 4.0 * R * (2.0 * G + 5.0 * H));
  if (r.bw_tc) {
    // one fused kernel per step: dh_rec = da_{t+1} Wh on tcgen05, backward pointwise in the epilogue
    if (last) {         // no recurrent gradient yet: pointwise only (dh_last rides in the recurrent slot)
      lstm_pointwise_bwd(cx, r.gates + (int64_t)t * R * G, cp, r.c + (int64_t)t * R * H, r.bw_dh_last, ext, nullptr, r.dc_carry, mk,
                         da_t, R, H);
      return;
    }
    bool ok = lstm_step_bwd_tc(cx, R, H, r.da + (int64_t)(t + 1) * R * G, Ws + (int64_t)D * G, r.gates + (int64_t)t * R * G, cp,
                               r.c + (int64_t)t * R * H, ext, r.dc_carry, mk, da_t);
    VD_REQUIRE(ok, VD_E_STATE, "lstm_step_bwd_tc refused a shape the engine routed to it");
    return;
  }
  const float* rec = last ? r.bw_dh_last : r.dh_rec;
  lstm_pointwise_bwd(cx, r.gates + (int64_t)t * R * G, cp, r.c + (int64_t)t * R * H, rec, ext, last ? r.bw_dc_last : nullptr,
                     r.dc_carry, mk, da_t, R, H);
  if (t > 0) gemm_tn((int)R, H, G, da_t, G, nullptr, Ws + (int64_t)D * G, G, r.dh_rec, H, 0.f, nullptr, 0);
}

void Engine::lstm_backward_end(LstmRun& r, float* dx_out, float* dh0_out, float* dc0_out) {
  const int H = r.H, D = r.D, G = 4 * r.H;
  const int64_t R = r.R, TR = (int64_t)r.T * R;
  const float* Ws = Wp(r.wseg);
  float* da = r.da;
  if (r.f16) VD_REQUIRE(!dh0_out && !dc0_out && !dx_out && !r.h0, VD_E_STATE, "fp16 BPTT: no initial-state / input gradients");
  if (dh0_out) gemm_tn((int)R, H, G, da, G, nullptr, Ws + (int64_t)D * G, G, dh0_out, H, 0.f, nullptr, 0);
  if (dc0_out) VD_CUDA_CHECK(cudaMemcpyAsync(dc0_out, r.dc_carry, (size_t)R * H * sizeof(float), cudaMemcpyDeviceToDevice, cx.stream));
  // accGradParameters
  float* dWs = dWp(r.wseg);
  const float* A = r.x ? r.x : Wp(0);
  const int64_t lda = r.x ? D : cfg.E;
  const bool pair16 = !r.f16 && r.h16 && r.da16 && H % 64 == 0 && TR - R >= 256;
  if (r.f16) {
    if (r.T > 1) {
      LaunchCtx::Scope sc(&cx, "gemm_wgrad", 2.0 * H * G * (double)(TR - R), 2.0 * (double)(TR - R) * (H + G));
      gemm_atb16(cx, H, G, TR - R, r.h16, H, r.da16 + R * G, G, dWs + (int64_t)D * G, G, r.scale2 + 1);
    }
  } else if (r.T > 1) {
    // persistent pair (enc_lstm.cu): its kernels left fp16 copies of h and da — the operands of the recurrence itself — so the weight
    // gradients contract those on the fp16 tensor-core path (fp32 accumulation) instead of the TF32 one
    if (pair16) {
      LaunchCtx::Scope sc(&cx, "gemm_wgrad", 2.0 * H * G * (double)(TR - R), 2.0 * (double)(TR - R) * (H + G));
      gemm_atb16(cx, H, G, TR - R, r.h16, H, r.da16 + R * G, G, dWs + (int64_t)D * G, G, nullptr);
    } else gemm_atb(H, G, TR - R, r.h, H, nullptr, da + R * G, G, dWs + (int64_t)D * G, G);
  }
  if (r.h0) gemm_atb(H, G, R, r.h0, H, nullptr, da, G, dWs + (int64_t)D * G, G);
  if (!r.x && tcmode()) {
    // Embedding-gathered input: every x_t is a row of the (V+1, E) table, so the three x-side gradients collapse
    // onto the table.  dP[v] = sum of da over the rows whose token is v (counting sort + balanced segmented sum,
    // HBM-bound) and then  dWx += E^T dP,  db += colsum(dP),  dEmb += dP Wx^T  are (V+1)-row contractions instead
    // of T*R-row ones (2 x 786 GFLOP -> 2 x 12 GFLOP for the 100 x 20-token option LSTM).
    VD_REQUIRE(D == cfg.E, VD_E_STATE, "gathered LSTM input must be the word embedding");
    const int V1 = cfg.V + 1;
    float* dP = arena.get<float>((int64_t)V1 * G);
    VD_CUDA_CHECK(cudaMemsetAsync(dP, 0, (size_t)V1 * G * sizeof(float), cx.stream));
    int32_t* scratch = arena.get<int32_t>(3 * (int64_t)V1);
    int32_t* perm = arena.get<int32_t>(TR);
    int32_t* stok = arena.get<int32_t>(TR);
    {
      LaunchCtx::Scope sc(&cx, "embed_grad_segsum", 0.0, (r.f16 ? 2.0 : 4.0) * TR * G);
      group_rows_by_token(cx, r.gather, TR, V1, scratch, perm, stok);
      if (r.f16) segsum_rows16(cx, r.da16, G, perm, stok, TR, dP, G, r.scale2 + 1);
      else segsum_rows(cx, da, G, perm, stok, TR, dP, G);
    }
    gemm_atb(D, G, V1, Wp(0), cfg.E, nullptr, dP, G, dWs, G);
    colsum_add(cx, dWp(r.wseg + 1), dP, V1, G, G);
    if (r.demb_out) gemm_tn(V1, D, G, dP, G, nullptr, Ws, G, r.demb_out, cfg.E, 0.f, nullptr, 0);   // folded in by the caller
    else gemm_tn(V1, D, G, dP, G, nullptr, Ws, G, dWp(0), cfg.E, 1.f, nullptr, 0);
    VD_REQUIRE(dx_out == nullptr, VD_E_STATE, "projected-space embedding gradient: caller must not ask for dx");
    return;
  }
  if (pair16 && r.x16 && !r.gather && D % 64 == 0) {
    LaunchCtx::Scope sc(&cx, "gemm_wgrad", 2.0 * D * G * (double)TR, 2.0 * (double)TR * (D + G));
    gemm_atb16(cx, D, G, TR, r.x16, D, r.da16, G, dWs, G, nullptr);
  } else gemm_atb(D, G, TR, A, lda, r.gather, da, G, dWs, G);
  colsum_add(cx, dWp(r.wseg + 1), da, TR, G, G);
  if (dx_out) gemm_tn((int)TR, D, G, da, G, nullptr, Ws, G, dx_out, D, 0.f, nullptr, 0);
}

void Engine::lstm_backward(LstmRun& r, const float* dh_all, const float* dh_last, const float* dc_last, float* dx_out,
                           float* dh0_out, float* dc0_out) {
  lstm_backward_begin(r, dh_all, dh_last, dc_last);
  for (int t = r.T - 1; t >= 0; --t) lstm_backward_step(r, t);
  lstm_backward_end(r, dx_out, dh0_out, dc0_out);
}

// BPTT wavefront of two stacked layers: layer-1 step t needs d(h1_t) = da2_t Wx2^T, produced per step on layer 2's
// stream, so the two recurrences again run one step apart.
void Engine::lstm_pair_backward(LstmRun& l1, LstmRun& l2, const float* dh_last2, const float* dc_last2, const float* dh_last1,
                                const float* dc_last1, float* dx1_out, cudaStream_t sa, cudaStream_t sb, cudaStream_t sc) {
  const int H = l2.H, G = 4 * l2.H;
  const int64_t R = l2.R;
  if (math_mode == VD_MATH_F16 && enc_persist_enabled() && l1.h16 && l2.h16 && l1.saved && l2.saved && l2.D == H && l1.T == l2.T &&
      enc_pair_shape_ok(R, H, cx.sm_count)) {
    // the forward of this pair ran as the persistent kernel: so does its BPTT (enc_lstm.cu::k_enc_pair_bwd) — one launch for
    // both layers and all time steps; the weight / input gradients follow as batched contractions over all T*R rows
    const int T = l2.T;
    const int64_t TR = (int64_t)T * R;
    cx.stream = sa;
    l1.da = arena.get<float>(TR * G); l2.da = arena.get<float>(TR * G);
    l1.da16 = arena.get<__half>(TR * G); l2.da16 = arena.get<__half>(TR * G);
    __half* Whb2 = arena.get<__half>((int64_t)H * G);
    __half* B1cat = arena.get<__half>((int64_t)H * 2 * G);
    cvt_f32_to_f16(cx, Whb2, G, Wp(l2.wseg) + (int64_t)l2.D * G, G, H, G);                 // Wh2: rows D2.. of (D2+H, 4H)
    cvt_f32_to_f16(cx, B1cat, 2 * G, Wp(l2.wseg), G, H, G);                                 // [Wx2 | ...]: rows 0..H-1 of layer 2
    cvt_f32_to_f16(cx, B1cat + G, 2 * G, Wp(l1.wseg) + (int64_t)l1.D * G, G, H, G);         // [... | Wh1]
    int* flags = arena.get<int>(enc_pair_bwd_flag_ints(T, R, H));
    float* dh1 = nullptr; float* dh2 = nullptr;
    if (enc_pair_gate_split(H)) { dh1 = arena.get<float>(TR * H); dh2 = arena.get<float>(TR * H); }
    {
      LaunchCtx::Scope sc2(&cx, "enc_pair_bwd", 2.0 * T * R * (double)H * 3.0 * G, 4.0 * T * R * (4.0 * G + 4.0 * H));
      enc_pair_backward(cx, T, R, H, B1cat, Whb2, l1.mask, l1.gates, l1.c, l2.gates, l2.c, dh_last1, dc_last1, dh_last2, dc_last2,
                        l1.da, l1.da16, l2.da, l2.da16, flags, dh1, dh2);
    }
    lstm_backward_end(l2, nullptr, nullptr, nullptr);
    lstm_backward_end(l1, dx1_out, nullptr, nullptr);
    return;
  }
  float* dx2 = arena.get<float>((int64_t)l2.T * R * l2.D);      // = gradient wrt layer-1 outputs, all steps
  const bool pipelined = tcmode() && H % 128 == 0 && sb != nullptr && sb != sa && R >= wavefront_min_rows();
  const bool three = pipelined && sc != nullptr && sc != sa && sc != sb && three_streams_enabled();
  cx.stream = sa;
  if (!pipelined) {
    lstm_backward(l2, nullptr, dh_last2, dc_last2, dx2, nullptr, nullptr);
    lstm_backward(l1, dx2, dh_last1, dc_last1, dx1_out, nullptr, nullptr);
    return;
  }
  const float* Wx2 = Wp(l2.wseg);            // rows 0..D2 of (D2+H, 4H): [N = D2, K = 4H]
  cudaEvent_t e0 = pool_event(0);
  VD_CUDA_CHECK(cudaEventRecord(e0, sa));
  VD_CUDA_CHECK(cudaStreamWaitEvent(sb, e0, 0));
  cx.stream = sb;
  lstm_backward_begin(l2, nullptr, dh_last2, dc_last2);
  cx.stream = sa;
  lstm_backward_begin(l1, dx2, dh_last1, dc_last1);
  for (int t = l2.T - 1; t >= 0; --t) {
    cx.stream = sb;
    lstm_backward_step(l2, t);
    cudaEvent_t e = pool_event(1 + t);
    if (three) {
      // d(h1_t) = da2_t Wx2 needs layer 2's step t only: on its own stream it runs beside layer 2's step t-1
      VD_CUDA_CHECK(cudaEventRecord(e, sb));
      VD_CUDA_CHECK(cudaStreamWaitEvent(sc, e, 0));
      cx.stream = sc;
    }
    gemm_tn((int)R, l2.D, G, l2.da + (int64_t)t * R * G, G, nullptr, Wx2, G, dx2 + (int64_t)t * R * l2.D, l2.D, 0.f, nullptr, 0);
    cudaEvent_t ex = three ? pool_event(1 + l2.T + t) : e;
    VD_CUDA_CHECK(cudaEventRecord(ex, cx.stream));
    VD_CUDA_CHECK(cudaStreamWaitEvent(sa, ex, 0));
    cx.stream = sa;
    lstm_backward_step(l1, t);
  }
  cx.stream = sb;
  lstm_backward_end(l2, nullptr, nullptr, nullptr);       // weight gradients of layer 2 (its dx was produced per step)
  VD_CUDA_CHECK(cudaEventRecord(e0, sb));
  cx.stream = sa;
  lstm_backward_end(l1, dx1_out, nullptr, nullptr);
  VD_CUDA_CHECK(cudaStreamWaitEvent(sa, e0, 0));
}

// ------------------------------------------------------------------------------------------------
// encoders
// ------------------------------------------------------------------------------------------------
static LstmRun make_run(int T, int64_t R, int D, int H, int wseg, const float* x, const int32_t* gather, const int32_t* mask) {
  LstmRun r; r.T = T; r.R = R; r.D = D; r.H = H; r.wseg = wseg; r.x = x; r.gather = gather; r.mask = mask;
  return r;
}

// ---- blocks shared by several encoder graphs -----------------------------------------------------------------------
// memory network over the history facts (mn-*.lua): MM -> MaskSoftMax (causal) -> MM -> Dropout -> Linear -> Tanh, residual query, Linear -> Tanh
void Engine::mn_block_fwd(const float* qin, const float* h3, float* out) {
  const int H = cfg.H, R = cfg.R, B = db.B;
  const int64_t N = db.N;
  const DropCfg d05 = dropcfg(0.5f);
  probs = arena.get<float>((int64_t)B * R * R);
  hAtt = arena.get<float>(N * H);
  mn_attention_fwd(cx, qin, h3, probs, hAtt, B, R, H);
  hAtt_d = arena.get<float>(N * H);
  dropout_apply(cx, hAtt_d, hAtt, N * H, d05, SITE_HATT);
  hAttTr = arena.get<float>(N * H);
  linear_fwd(seg("mn.fact.weight"), hAtt_d, N, hAttTr, 1);
  sum1 = arena.get<float>(N * H);
  add_out(cx, sum1, hAttTr, qin, N * H);
  linear_fwd(seg("mn.query.weight"), sum1, N, out, 1);
}
// dout = gradient wrt `out` (post-tanh); dqin / dh3 receive the gradients wrt the query input and the facts
void Engine::mn_block_bwd(const float* dout, const float* out, float* dqin, float* dh3) {
  const int H = cfg.H, R = cfg.R, B = db.B;
  const int64_t N = db.N;
  const DropCfg d05 = dropcfg(0.5f);
  float* dpre = arena.get<float>(N * H);
  tanh_bwd(cx, dpre, dout, out, N * H);
  float* dsum1 = arena.get<float>(N * H);
  linear_bwd(seg("mn.query.weight"), sum1, dpre, N, dsum1, 0.f);
  tanh_bwd(cx, dpre, dsum1, hAttTr, N * H);
  float* dhAtt = arena.get<float>(N * H);
  linear_bwd(seg("mn.fact.weight"), hAtt_d, dpre, N, dhAtt, 0.f);
  dropout_apply(cx, dhAtt, dhAtt, N * H, d05, SITE_HATT);
  mn_attention_bwd(cx, mn_query_in, hist2.h_last(), probs, dhAtt, dqin, dh3, B, R, H);
  add_inplace(cx, dqin, dsum1, N * H);
}
// SAN (mn-att-ques-im-hist.lua:67-106, lf-att-ques-im-hist.lua:43-86): tanh(Linear(img)) is computed once per dialog; Dropout then
// acts on the repeated tensor; hops of {img_common + ques_common -> tanh -> dropout -> score -> softmax -> weighted sum + u}
void Engine::san_block_fwd(const float* u0) {
  const int H = cfg.H, R = cfg.R, B = db.B;
  const int64_t N = db.N;
  const int P = cfg.S * cfg.S, Cm = cfg.Cm;
  const DropCfg d05 = dropcfg(0.5f);
  t_img = arena.get<float>((int64_t)B * P * H);
  wait_img();
  linear_fwd(seg("san.img.weight"), db.img, (int64_t)B * P, t_img, 1);
  img_tr = arena.get<float>(N * P * H);
  san_expand_dropout(cx, img_tr, t_img, B, R, P, H, d05, SITE_IMG_TR);
  img_common.assign(cfg.hops, nullptr); ques_common.assign(cfg.hops, nullptr);
  sc.assign(cfg.hops, nullptr); pr.assign(cfg.hops, nullptr); u_hop.assign(cfg.hops + 1, nullptr);
  u_hop[0] = const_cast<float*>(u0);
  for (int hop = 0; hop < cfg.hops; ++hop) {
    std::string pre = "san.hop" + std::to_string(hop + 1) + ".";
    int w_ic = seg((pre + "img_common.weight").c_str()), w_qc = seg((pre + "ques_common.weight").c_str()),
        w_s = seg((pre + "score.weight").c_str());
    img_common[hop] = arena.get<float>(N * P * Cm);
    linear_fwd(w_ic, img_tr, N * P, img_common[hop], 0);
    ques_common[hop] = arena.get<float>(N * Cm);
    linear_fwd(w_qc, u_hop[hop], N, ques_common[hop], 0);
    sc[hop] = arena.get<float>(N * P);
    san_score_fwd(cx, img_common[hop], ques_common[hop], Wp(w_s), Wp(w_s + 1), sc[hop], N, P, Cm, d05, SITE_HOP0 + hop);
    pr[hop] = arena.get<float>(N * P);
    u_hop[hop + 1] = arena.get<float>(N * H);
    san_softmax_att_fwd(cx, sc[hop], pr[hop], img_tr, u_hop[hop], u_hop[hop + 1], N, P, H);
  }
  u_d = arena.get<float>(N * H);
  dropout_apply(cx, u_d, u_hop[cfg.hops], N * H, d05, SITE_U_OUT);
  linear_fwd(seg("san.out.weight"), u_d, N, encOut, 1);
}
void Engine::san_block_bwd(const float* dEnc, float* du) {
  const int H = cfg.H, R = cfg.R, B = db.B;
  const int64_t N = db.N;
  const int P = cfg.S * cfg.S, Cm = cfg.Cm;
  const DropCfg d05 = dropcfg(0.5f);
  float* dpre = arena.get<float>(N * H);
  tanh_bwd(cx, dpre, dEnc, encOut, N * H);
  linear_bwd(seg("san.out.weight"), u_d, dpre, N, du, 0.f);
  dropout_apply(cx, du, du, N * H, d05, SITE_U_OUT);
  float* dimg_tr = arena.get<float>(N * P * H);
  float* ds = arena.get<float>(N * P);
  float* dic = arena.get<float>(N * P * Cm);
  float* dqc = arena.get<float>(N * Cm);
  for (int hop = cfg.hops - 1; hop >= 0; --hop) {
    std::string pre = "san.hop" + std::to_string(hop + 1) + ".";
    int w_ic = seg((pre + "img_common.weight").c_str()), w_qc = seg((pre + "ques_common.weight").c_str()),
        w_s = seg((pre + "score.weight").c_str());
    if (hop == cfg.hops - 1) {
      san_att_bwd(cx, du, pr[hop], img_tr, ds, dimg_tr, N, P, H);
    } else {
      float* tmp = arena.get<float>(N * P * H);
      san_att_bwd(cx, du, pr[hop], img_tr, ds, tmp, N, P, H);
      add_inplace(cx, dimg_tr, tmp, N * P * H);
    }
    san_score_bwd(cx, ds, img_common[hop], ques_common[hop], Wp(w_s), dic, dqc, dWp(w_s), dWp(w_s + 1), N, P, Cm, d05,
                  SITE_HOP0 + hop);
    linear_bwd(w_ic, img_tr, dic, N * P, dimg_tr, 1.f);
    linear_bwd(w_qc, u_hop[hop], dqc, N, du, 1.f);          // du now = grad wrt u_hop[hop]
  }
  float* dt_pre = arena.get<float>((int64_t)B * P * H);
  san_collapse_bwd(cx, dimg_tr, t_img, dt_pre, B, R, P, H, d05, SITE_IMG_TR);
  linear_bwd(seg("san.img.weight"), db.img, dt_pre, (int64_t)B * P, nullptr, 0.f);
  release_img();                                   // last reader of this step's image staging buffer
}

void Engine::encoder_forward(const vd_batch* b) {
  VD_CUDA_CHECK(cudaSetDevice(cfg.gpuid));
  if (side_active) { cudaStreamSynchronize(side_stream); side_active = false; }   // only after an aborted call
  cx.stream = main_stream;
  if (opt_fwd_pending || opt_bwd_pending) {       // a previous call sequence was abandoned half-way: drain before reuse
    VD_CUDA_CHECK(cudaStreamSynchronize(opt_stream));
    opt_fwd_pending = opt_bwd_pending = false;
  }
  arena.reset();
  have_fwd = false;
  stage_batch(b);
  save_acts = training != 0;
  refresh_shadows();
  if (opt_overlap && cfg.dec == DEC_DISC && db.options && db.To > 0) options_forward_async();
  conn_dh_l1 = conn_dc_l1 = conn_dc_l2 = nullptr;
  gen_h0[0] = gen_h0[1] = gen_c0[0] = gen_c0[1] = nullptr;
  const int E = cfg.E, H = cfg.H, R = cfg.R;
  const int64_t N = db.N;
  const int B = db.B;
  const DropCfg d05 = dropcfg(0.5f), dp = dropcfg(cfg.dropout), dnone = dropcfg(0.f);
  // time-major ids: the view(-1,T):t() of model.lua:255-257,274-278
  ids_q = arena.get<int32_t>(N * db.Tq);
  transpose_ids(cx, db.ques, ids_q, N, db.Tq);
  if (cfg.useHist) {
    ids_h = arena.get<int32_t>(N * db.Th);
    transpose_ids(cx, db.hist, ids_h, N, db.Th);
  }
  const bool embdrop = cfg.embdrop;                        // mn-*.lua / lf-att-*.lua: Dropout(0.5) on both embeddings
  // history branch
  if (cfg.useHist) {
    fork_side();
    xh = arena.get<float>(N * db.Th * E);
    embed_rows(cx, xh, Wp(0), ids_h, N * db.Th, E, embdrop ? d05 : dnone, SITE_HEMBED);
    hist1 = make_run(db.Th, N, E, H, seg("hist.lstm1.weight"), xh, nullptr, ids_h);
    hist2 = make_run(db.Th, N, H, H, seg("hist.lstm2.weight"), nullptr, nullptr, ids_h);
  }
  // The 2nd layer consumes every step of the 1st, so encoder LSTMs always keep all steps (T*N*H is small).
  auto run_two = [&](LstmRun& l1, LstmRun& l2, cudaStream_t sa, cudaStream_t sb, cudaStream_t sc) { lstm_pair_forward(l1, l2, sa, sb, sc); };
  // The history and question LSTM chains are independent until the fusion/attention stage: the history chain runs
  // on the side stream (its tiny per-step kernels are latency-bound; overlapping the two chains hides half of it).
  // (the persistent pair kernels of VD_MATH_F16 are flag-chained grids that want every SM: two of them must never be
  //  co-scheduled — neither could become fully resident — so in that mode both pairs run on the main stream, in order)
  const bool serial_pairs = math_mode == VD_MATH_F16 && enc_persist_enabled();
  if (cfg.useHist) {
    if (serial_pairs) { join_side(); run_two(hist1, hist2, main_stream, main2_stream, main3_stream); }
    else { fork_side(); run_two(hist1, hist2, side_stream, side2_stream, side3_stream); back_to_main(); }
  }
  // question branch
  xq = arena.get<float>(N * db.Tq * E);
  embed_rows(cx, xq, Wp(0), ids_q, N * db.Tq, E, embdrop ? d05 : dnone, SITE_QEMBED);
  if (cfg.img_in_q) {
    // hre(a)-ques-im-hist.lua:41-55: [Dropout(0.5) ->] Linear(F,IE) on the 10x repeated fc7, MaskTime, JoinTable(-1)
    img_d = arena.get<float>(N * cfg.F);
    wait_img();
    repeat_rows(cx, img_d, db.img, B, R, cfg.F);
    if (cfg.img_drop) dropout_apply(cx, img_d, img_d, N * cfg.F, d05, SITE_IMG_FC7);
    img_e = arena.get<float>(N * cfg.IE);
    linear_fwd(seg("img.embed.weight"), img_d, N, img_e, 0);
    qi_in = arena.get<float>(N * db.Tq * (E + cfg.IE));
    masktime_concat_fwd(cx, qi_in, xq, img_e, ids_q, db.Tq, N, E, cfg.IE);
    ques1 = make_run(db.Tq, N, E + cfg.IE, H, seg("ques.lstm1.weight"), qi_in, nullptr, ids_q);
  } else {
    ques1 = make_run(db.Tq, N, E, H, seg("ques.lstm1.weight"), xq, nullptr, ids_q);
  }
  ques2 = make_run(db.Tq, N, H, H, seg("ques.lstm2.weight"), nullptr, nullptr, ids_q);
  run_two(ques1, ques2, main_stream, main2_stream, main3_stream);
  join_side();
  const float* q3 = ques2.h_last();
  const float* h3 = cfg.useHist ? hist2.h_last() : nullptr;
  encOut = arena.get<float>(N * H);

  if (cfg.fam_lf) {
    // lf-ques.lua:29-33 / lf-ques-im.lua / lf-ques-hist.lua / lf-ques-im-hist.lua:49-59: JoinTable [q | img | h] -> Dropout -> Linear
    // -> Tanh.  lf-att-ques-im-hist.lua:41: Tanh(Linear([q | h])) with no dropout, then the SAN block on pool5.
    const bool fc7 = cfg.useIm && !cfg.san;
    joinK = H + (fc7 ? cfg.F : 0) + (cfg.useHist ? H : 0);
    join_d = arena.get<float>(N * joinK);
    copy_cols(cx, join_d, joinK, q3, H, N, H);
    if (fc7) {                                          // image repeated per round (model.lua:267-269)
      float* tmp = arena.get<float>(N * cfg.F);
      wait_img();
      repeat_rows(cx, tmp, db.img, B, R, cfg.F);
      copy_cols(cx, join_d + H, joinK, tmp, cfg.F, N, cfg.F);
    }
    if (cfg.useHist) copy_cols(cx, join_d + H + (fc7 ? cfg.F : 0), joinK, h3, H, N, H);
    if (!cfg.san) {
      dropout_apply(cx, join_d, join_d, N * joinK, dp, SITE_FUSION);
      linear_fwd(seg("fusion.weight"), join_d, N, encOut, 1);
    } else {
      qh2 = arena.get<float>(N * H);
      linear_fwd(seg("fusion.weight"), join_d, N, qh2, 1);
      san_block_fwd(qh2);
    }
  } else if (cfg.fam_hre) {
    // hrea-ques-im-hist.lua:89-137 (attention over the rounds, [att | q]) / hre-ques-*.lua ([q | h]) -> dialog-level LSTM
    jt = arena.get<float>(N * 2 * H);
    float* j = arena.get<float>(N * 2 * H);
    if (cfg.hre_att) {
      sq = arena.get<float>(N); sh = arena.get<float>(N);
      int sq_w = seg("att.q.weight"), sh_w = seg("att.h.weight");
      rowdot_fwd(cx, sq, q3, Wp(sq_w), Wp(sq_w + 1), N, H);
      rowdot_fwd(cx, sh, h3, Wp(sh_w), Wp(sh_w + 1), N, H);
      probs = arena.get<float>((int64_t)B * R * R);
      att = arena.get<float>(N * H);
      hrea_attention_fwd(cx, sq, sh, h3, probs, att, B, R, H);
      copy_cols(cx, j, 2 * H, att, H, N, H);
      copy_cols(cx, j + H, 2 * H, q3, H, N, H);
    } else {
      copy_cols(cx, j, 2 * H, q3, H, N, H);
      copy_cols(cx, j + H, 2 * H, h3, H, N, H);
    }
    // View(-1,10,2H), Transpose(1,2): row (r,b) <- row (b,r), a strided 2-D copy per round
    for (int rr = 0; rr < R; ++rr)
      copy_cols(cx, jt + (int64_t)rr * B * 2 * H, 2 * H, j + (int64_t)rr * 2 * H, (int64_t)R * 2 * H, B, 2 * H);
    dialog = make_run(R, B, 2 * H, H, seg("dialog.lstm.weight"), jt, nullptr, nullptr);
    lstm_forward(dialog, true);
    for (int rr = 0; rr < R; ++rr)
      copy_cols(cx, encOut + (int64_t)rr * H, (int64_t)R * H, dialog.h + (int64_t)rr * B * H, H, B, H);
  } else {
    // mn-ques-hist.lua / mn-ques-im-hist.lua / mn-att-ques-im-hist.lua:48-106
    mn_query_in = q3;
    if (cfg.mn_qi) {                                    // mn-ques-im-hist.lua: qi_proj = Tanh(Linear([q | fc7]))
      qi_join = arena.get<float>(N * (H + cfg.F));
      copy_cols(cx, qi_join, H + cfg.F, q3, H, N, H);
      float* tmp = arena.get<float>(N * cfg.F);
      wait_img();
      repeat_rows(cx, tmp, db.img, B, R, cfg.F);
      copy_cols(cx, qi_join + H, H + cfg.F, tmp, cfg.F, N, cfg.F);
      qi_proj = arena.get<float>(N * H);
      linear_fwd(seg("mn.qi.weight"), qi_join, N, qi_proj, 1);
      mn_query_in = qi_proj;
    }
    if (cfg.san) {
      qh2 = arena.get<float>(N * H);
      mn_block_fwd(mn_query_in, h3, qh2);
      san_block_fwd(qh2);
    } else {
      mn_block_fwd(mn_query_in, h3, encOut);
    }
  }
  wait_img();                // (an encoder that never reads the image still has to retire the copy before the buffer is reused)
  release_img();             // forward-only callers never reach the backward's release; a later one simply overrides this
  have_fwd = true;
}

void Engine::encoder_backward(const float* dEnc) {
  VD_REQUIRE(have_fwd && save_acts, VD_E_STATE, "encoder_backward: no training-mode forward to back-propagate");
  VD_REQUIRE(dEnc != nullptr, VD_E_BADARG, "gradEncOut is null");
  const int E = cfg.E, H = cfg.H, R = cfg.R;
  const int64_t N = db.N;
  const int B = db.B;
  const DropCfg d05 = dropcfg(0.5f), dp = dropcfg(cfg.dropout), dnone = dropcfg(0.f);
  float* dq3 = arena.get<float>(N * H);
  float* dh3 = cfg.useHist ? arena.get<float>(N * H) : nullptr;
  float* dpre = arena.get<float>(N * H);
  // gradient sync, bucket 0: the decoder's own weights are final once decoder_backward is enqueued (gen: dec.lstm1/2 +
  // dec.out, 20 MB at V = 10k; disc without the option stream: opt.lstm) — their all-reduce overlaps the whole encoder BPTT
  if (cfg.dec == DEC_GEN) reduce_segments(seg("dec.lstm1.weight"), (int)lay.segs.size() - 1, main_stream);
  else if (!opt_bwd_pending) reduce_segments(seg("opt.lstm.weight"), (int)lay.segs.size() - 1, main_stream);

  if (cfg.fam_lf) {
    const bool fc7 = cfg.useIm && !cfg.san;
    float* dj = arena.get<float>(N * joinK);
    if (!cfg.san) {
      tanh_bwd(cx, dpre, dEnc, encOut, N * H);
      linear_bwd(seg("fusion.weight"), join_d, dpre, N, dj, 0.f);
      dropout_apply(cx, dj, dj, N * joinK, dp, SITE_FUSION);
    } else {
      float* dqh = arena.get<float>(N * H);
      san_block_bwd(dEnc, dqh);
      tanh_bwd(cx, dpre, dqh, qh2, N * H);
      linear_bwd(seg("fusion.weight"), join_d, dpre, N, dj, 0.f);
    }
    copy_cols(cx, dq3, H, dj, joinK, N, H);
    if (cfg.useHist) copy_cols(cx, dh3, H, dj + H + (fc7 ? cfg.F : 0), joinK, N, H);
  } else if (cfg.fam_hre) {
    const float* q3 = ques2.h_last();
    const float* h3 = hist2.h_last();
    // un-permute the gradient (n = b*R + r) -> (r,b), BPTT over rounds
    float* dd = arena.get<float>(N * H);
    for (int rr = 0; rr < R; ++rr)
      copy_cols(cx, dd + (int64_t)rr * B * H, H, dEnc + (int64_t)rr * H, (int64_t)R * H, B, H);
    float* djt = arena.get<float>(N * 2 * H);
    lstm_backward(dialog, dd, nullptr, nullptr, djt, nullptr, nullptr);
    float* dj = arena.get<float>(N * 2 * H);
    for (int rr = 0; rr < R; ++rr)
      copy_cols(cx, dj + (int64_t)rr * 2 * H, (int64_t)R * 2 * H, djt + (int64_t)rr * B * 2 * H, 2 * H, B, 2 * H);
    if (cfg.hre_att) {
      float* datt = arena.get<float>(N * H);
      copy_cols(cx, datt, H, dj, 2 * H, N, H);
      copy_cols(cx, dq3, H, dj + H, 2 * H, N, H);
      float* dsq = arena.get<float>(N); float* dsh = arena.get<float>(N);
      hrea_attention_bwd(cx, sq, sh, h3, probs, datt, dsq, dsh, dh3, B, R, H);
      int sq_w = seg("att.q.weight"), sh_w = seg("att.h.weight");
      rowdot_bwd(cx, dsq, q3, Wp(sq_w), dq3, 1, dWp(sq_w), dWp(sq_w + 1), N, H);
      rowdot_bwd(cx, dsh, h3, Wp(sh_w), dh3, 1, dWp(sh_w), dWp(sh_w + 1), N, H);
    } else {
      copy_cols(cx, dq3, H, dj, 2 * H, N, H);
      copy_cols(cx, dh3, H, dj + H, 2 * H, N, H);
    }
  } else {
    // memory-network family: [SAN ->] memory block -> [qi projection]
    float* dqin = cfg.mn_qi ? arena.get<float>(N * H) : dq3;
    if (cfg.san) {
      float* du = arena.get<float>(N * H);
      san_block_bwd(dEnc, du);
      mn_block_bwd(du, qh2, dqin, dh3);
    } else {
      mn_block_bwd(dEnc, encOut, dqin, dh3);
    }
    if (cfg.mn_qi) {
      tanh_bwd(cx, dpre, dqin, qi_proj, N * H);
      float* dqj = arena.get<float>(N * (H + cfg.F));
      linear_bwd(seg("mn.qi.weight"), qi_join, dpre, N, dqj, 0.f);
      copy_cols(cx, dq3, H, dqj, H + cfg.F, N, H);
    }
    // gradient sync, bucket 1: every non-recurrent layer of the encoder is final here, before the LSTM BPTTs start
    if (cfg.san) reduce_segments(seg(cfg.mn_qi ? "mn.qi.weight" : "mn.fact.weight"), seg("san.out.weight") + 1, main_stream);
  }

  // question LSTMs (+ gradients handed back by the gen decoder, gen.lua:45-60); the history chain's BPTT runs
  // concurrently on the side stream (disjoint weight segments; the shared embedding gradient is atomics-only)
  const bool embdrop = cfg.embdrop;
  // (persistent pair kernels must not be co-scheduled — see encoder_forward: in that mode both BPTTs run on the main stream)
  const bool serial_pairs = math_mode == VD_MATH_F16 && enc_persist_enabled() && hist1.h16 != nullptr;
  if (cfg.useHist) {
    if (serial_pairs) { join_side(); cx.stream = main_stream; } else fork_side();
    cudaStream_t hs = serial_pairs ? main_stream : side_stream;
    float* dx1 = arena.get<float>(N * db.Th * E);
    lstm_pair_backward(hist1, hist2, dh3, nullptr, nullptr, nullptr, dx1, hs, serial_pairs ? main2_stream : side2_stream,
                       serial_pairs ? main3_stream : side3_stream);
    reduce_segments(seg("hist.lstm1.weight"), seg("hist.lstm2.weight") + 1, hs);                // bucket 2
    embed_scatter_add(cx, dWp(0), dx1, E, ids_h, N * db.Th, E, embdrop ? d05 : dnone, SITE_HEMBED);
    back_to_main();
  }
  {
    int D1 = ques1.D;
    float* dx1 = arena.get<float>(N * db.Tq * D1);
    lstm_pair_backward(ques1, ques2, dq3, conn_dc_l2, conn_dh_l1, conn_dc_l1, dx1, main_stream, main2_stream, main3_stream);
    reduce_segments(seg("ques.lstm1.weight"), seg("ques.lstm2.weight") + 1, main_stream);       // bucket 3
    embed_scatter_add(cx, dWp(0), dx1, D1, ids_q, N * db.Tq, E, embdrop ? d05 : dnone, SITE_QEMBED);
    if (cfg.img_in_q) {
      float* die = arena.get<float>(N * cfg.IE);
      masktime_bwd(cx, dx1, D1, E, ids_q, die, db.Tq, N, cfg.IE);
      linear_bwd(seg("img.embed.weight"), img_d, die, N, nullptr, 0.f);
    }
  }
  join_side();
  // what is left — the option LSTM's weights (final when the option stream ends, i.e. now), the word embedding, which every
  // branch writes, and the small layers not covered above — goes in clamp_adam_step
  // (opt.lstm, final at the same moment, goes out with the embedding in the grouped launch of reduce_remaining)
  join_options_backward();
}

void Engine::fork_side() {
  if (!side_active) {
    VD_CUDA_CHECK(cudaEventRecord(ev_fork, main_stream));
    VD_CUDA_CHECK(cudaStreamWaitEvent(side_stream, ev_fork, 0));
    side_active = true;
  }
  cx.stream = side_stream;
}
void Engine::back_to_main() { cx.stream = main_stream; }
void Engine::join_side() {
  cx.stream = main_stream;
  if (side_active) {
    VD_CUDA_CHECK(cudaEventRecord(ev_join, side_stream));
    VD_CUDA_CHECK(cudaStreamWaitEvent(main_stream, ev_join, 0));
    side_active = false;
  }
}

// ------------------------------------------------------------------------------------------------
// decoders + criterions
// ------------------------------------------------------------------------------------------------
void Engine::forward_connect() {
  // decoders/gen.lua:30-42; decoders/disc.lua:35 is a no-op
  if (cfg.dec != DEC_GEN) return;
  VD_REQUIRE(have_fwd, VD_E_STATE, "forward_connect before encoder_forward");
  gen_h0[0] = gen_h0[1] = gen_c0[0] = gen_c0[1] = nullptr;
  if (cfg.rnn_layers) {                              // encoders with .rnnLayers
    gen_h0[0] = ques1.h_last(); gen_c0[0] = ques1.c_last();
    gen_c0[1] = ques2.c_last();
  }
  gen_h0[1] = encOut;
}

// disc.lua:4-20: the option LSTM over all N*100 candidate answers.  Depends only on the batch and the weights, so it
// is enqueued on opt_stream before the encoder's kernels and meets the encoder output at disc_scores_fwd.
void Engine::options_forward_async() {
  const int64_t Ro = db.N * cfg.K;
  cudaStream_t s = opt_overlap ? opt_stream : main_stream;
  if (opt_overlap) {
    VD_CUDA_CHECK(cudaEventRecord(ev_opt_fork, main_stream));          // batch staged, shadows refreshed
    VD_CUDA_CHECK(cudaStreamWaitEvent(opt_stream, ev_opt_fork, 0));
  }
  cudaStream_t prev = cx.stream;
  cx.stream = s;
  if (opt_overlap) {                                                // leave SMs to the encoder's concurrent chains
    static int fwd_reserve = -1;                                    // VD_OPT_RESERVE_FWD: A/B knob, default = the common reserve
    if (fwd_reserve < 0) { const char* e = getenv("VD_OPT_RESERVE_FWD"); fwd_reserve = e ? (atoi(e) & ~1) : 1 << 20; }
    // with the persistent encoder forward (VD_MATH_F16) nothing latency-bound runs beside the option stream's forward: no reserve
    const int dflt = (math_mode == VD_MATH_F16 && enc_persist_enabled()) ? 0 : opt_reserve_sms;
    cx.sm_budget = cx.sm_count - (fwd_reserve == 1 << 20 ? dflt : std::min(fwd_reserve, cx.sm_count - 16));
  }
  ids_o = arena.get<int32_t>(Ro * db.To);
  transpose_ids(cx, db.options, ids_o, Ro, db.To);
  opt = make_run(db.To, Ro, cfg.E, cfg.H, seg("opt.lstm.weight"), nullptr, ids_o, nullptr);   // disc.lua:4-5: no maskzero
  lstm_forward(opt, save_acts);
  VD_CUDA_CHECK(cudaEventRecord(ev_opt_done, s));
  cx.stream = prev;
  cx.sm_budget = 0;
  opt_fwd_pending = true;
}

void Engine::join_options_backward() {
  if (!opt_bwd_pending) return;
  opt_bwd_pending = false;
  cudaStream_t prev = cx.stream;
  cx.stream = main_stream;
  VD_CUDA_CHECK(cudaStreamWaitEvent(main_stream, ev_opt_done, 0));
  if (opt_demb) add_inplace(cx, dWp(0), opt_demb, (int64_t)(cfg.V + 1) * cfg.E);
  opt_demb = nullptr;
  cx.stream = prev;
}

void Engine::decoder_forward() {
  VD_REQUIRE(have_fwd, VD_E_STATE, "decoder_forward before encoder_forward");
  const int E = cfg.E, H = cfg.H, K = cfg.K;
  const int64_t N = db.N;
  if (cfg.dec == DEC_DISC) {
    VD_REQUIRE(db.options && db.To > 0, VD_E_SHAPE, "disc decoder: options / To missing");
    if (!opt_fwd_pending) options_forward_async();          // normally started by encoder_forward already
    VD_CUDA_CHECK(cudaStreamWaitEvent(main_stream, ev_opt_done, 0));
    opt_fwd_pending = false;
    cx.stream = main_stream;
    scores = arena.get<float>(N * K);
    disc_scores_fwd(cx, opt.h_last(), encOut, scores, N, K, H);
  } else {
    VD_REQUIRE(db.answer_in && db.Ta > 0, VD_E_SHAPE, "gen decoder: answer_in / Ta missing");
    forward_connect();
    ids_ai = arena.get<int32_t>(N * db.Ta);
    transpose_ids(cx, db.answer_in, ids_ai, N, db.Ta);
    float* xa = arena.get<float>(N * db.Ta * E);
    embed_rows(cx, xa, Wp(0), ids_ai, N * db.Ta, E, dropcfg(0.f), 0);
    dec1 = make_run(db.Ta, N, E, H, seg("dec.lstm1.weight"), xa, nullptr, ids_ai);
    dec1.h0 = gen_h0[0]; dec1.c0 = gen_c0[0];
    lstm_forward(dec1, true);
    dec2 = make_run(db.Ta, N, H, H, seg("dec.lstm2.weight"), dec1.h, nullptr, ids_ai);
    dec2.h0 = gen_h0[1]; dec2.c0 = gen_c0[1];
    lstm_forward(dec2, true);
    fused_vocab_fwd = false;
    const int64_t rows = N * db.Ta;
    const int wo = seg("dec.out.weight");
    if (!want_logp && tcmode() && db.answer_out) {
      // whole-step entry point: nobody reads decOut, only the criterion does — keep the logits on chip
      ids_ao = arena.get<int32_t>(rows);
      transpose_ids(cx, db.answer_out, ids_ao, N, db.Ta);
      voc_nparts = vocab_lse_nparts(cfg.V);
      voc_pm = arena.get<float>(rows * voc_nparts); voc_ps = arena.get<float>(rows * voc_nparts);
      voc_tl = arena.get<float>(rows); voc_lse = arena.get<float>(rows);
      LaunchCtx::Scope sc(&cx, "vocab_lse", 2.0 * rows * cfg.V * H, 4.0 * (rows * (double)H + (double)cfg.V * H + 2.0 * rows * voc_nparts));
      fused_vocab_fwd = vocab_lse_tc(cx, (int)rows, cfg.V, H, dec2.h, H, Wp(wo), H, Wp(wo + 1), ids_ao, voc_pm, voc_ps, voc_tl);
    }
    if (!fused_vocab_fwd) {
      logp = arena.get<float>(rows * cfg.V);
      linear_fwd(wo, dec2.h, rows, logp, 0);
      logsoftmax_rows(cx, logp, ids_ai, rows, cfg.V);            // gen.lua:23-24 (MaskZero)
    } else {
      logp = nullptr;
    }
  }
}

float Engine::criterion_forward() {
  const int64_t N = db.N;
  if (cfg.dec == DEC_DISC) {
    VD_REQUIRE(scores && db.answer_ind, VD_E_STATE, "criterion: decoder_forward / answer_ind missing");
    row_loss = arena.get<float>(N);
    xent_fwd(cx, scores, db.answer_ind, row_loss, N, cfg.K);
    reduce_sum(cx, row_loss, scalars_dev, N, 1.f / (float)N);        // CrossEntropyCriterion: mean
  } else {
    VD_REQUIRE((logp || fused_vocab_fwd) && db.answer_out, VD_E_STATE, "criterion: decoder_forward / answer_out missing");
    row_loss = arena.get<float>(N * db.Ta);
    if (fused_vocab_fwd) {
      vocab_lse_finish(cx, voc_pm, voc_ps, voc_nparts, voc_tl, ids_ao, ids_ai, voc_lse, row_loss, -1.f, 0, N * db.Ta);
    } else {
      ids_ao = arena.get<int32_t>(N * db.Ta);
      transpose_ids(cx, db.answer_out, ids_ao, N, db.Ta);
      nll_fwd(cx, logp, ids_ao, ids_ai, row_loss, N * db.Ta, cfg.V);
    }
    reduce_sum(cx, row_loss, scalars_dev, N * db.Ta, 1.f);            // ClassNLL sizeAverage=false
  }
  float loss = 0.f;
  VD_CUDA_CHECK(cudaMemcpyAsync(&loss, scalars_dev, sizeof(float), cudaMemcpyDeviceToHost, cx.stream));
  VD_CUDA_CHECK(cudaStreamSynchronize(cx.stream));                    // the reference reads curLoss here too (model.lua:330)
  return loss;
}

void Engine::criterion_backward() {
  const int64_t N = db.N;
  if (cfg.dec == DEC_DISC) {
    VD_REQUIRE(scores && db.answer_ind, VD_E_STATE, "criterion_backward: decoder_forward / answer_ind missing");
    dscores = arena.get<float>(N * cfg.K);
    xent_bwd(cx, scores, db.answer_ind, dscores, N, cfg.K);
  } else {
    VD_REQUIRE((logp || fused_vocab_fwd) && ids_ao, VD_E_STATE, "criterion_backward before criterion_forward");
    dlogits = arena.get<float>(N * db.Ta * cfg.V);
    if (fused_vocab_fwd) {
      const int wo = seg("dec.out.weight");
      LaunchCtx::Scope sc(&cx, "vocab_dlogits", 2.0 * N * db.Ta * cfg.V * cfg.H, 4.0 * N * db.Ta * (double)cfg.V);
      bool ok = vocab_dlogits_tc(cx, (int)(N * db.Ta), cfg.V, cfg.H, dec2.h, cfg.H, Wp(wo), cfg.H, Wp(wo + 1), ids_ao, ids_ai, voc_lse,
                                 dlogits, cfg.V);
      VD_REQUIRE(ok, VD_E_STATE, "vocab_dlogits_tc refused a shape vocab_lse_tc took");
    } else {
      nll_bwd(cx, logp, ids_ao, ids_ai, dlogits, N * db.Ta, cfg.V);
    }
  }
}

void Engine::decoder_backward() {
  VD_REQUIRE(save_acts, VD_E_STATE, "decoder_backward needs a training-mode forward");
  if (world > 1)
    for (char c : seg_reduced)
      VD_REQUIRE(!c, VD_E_STATE, "world > 1: vd_zero_grad is required between backward passes (gradients already all-reduced)");
  const int E = cfg.E, H = cfg.H, K = cfg.K;
  const int64_t N = db.N;
  dEncFromDec = arena.get<float>(N * H);
  if (cfg.dec == DEC_DISC) {
    VD_REQUIRE(dscores, VD_E_STATE, "decoder_backward before criterion_backward");
    const int64_t Ro = N * K;
    float* dfeat = arena.get<float>(Ro * H);
    disc_scores_bwd(cx, dscores, opt.h_last(), encOut, dfeat, dEncFromDec, N, K, H);
    // The option BPTT feeds only opt.lstm's weights and the word embedding: it runs on opt_stream while the caller
    // goes on to the encoder's backward; join_options_backward() (encoder_backward / any gradient consumer) waits.
    cudaStream_t prev = cx.stream;
    if (opt_overlap) {
      VD_CUDA_CHECK(cudaEventRecord(ev_opt_fork, cx.stream));
      VD_CUDA_CHECK(cudaStreamWaitEvent(opt_stream, ev_opt_fork, 0));
      cx.stream = opt_stream;
      cx.sm_budget = cx.sm_count - opt_reserve_sms;
    }
    if (tcmode()) {
      // embedding gradient in projected space; written to a private (V+1,E) buffer when overlapped, because the
      // encoder's embedding gradients accumulate into dW(wordEmbed) with atomics at the same time
      opt.demb_out = opt_overlap ? (opt_demb = arena.get<float>((int64_t)(cfg.V + 1) * E)) : nullptr;
      lstm_backward(opt, nullptr, dfeat, nullptr, nullptr, nullptr, nullptr);
    } else {
      float* dx = arena.get<float>(Ro * db.To * E);
      lstm_backward(opt, nullptr, dfeat, nullptr, dx, nullptr, nullptr);
      embed_scatter_add(cx, dWp(0), dx, E, ids_o, Ro * db.To, E, dropcfg(0.f), 0);   // atomics: safe next to the encoder's
    }
    if (opt_overlap) {
      VD_CUDA_CHECK(cudaEventRecord(ev_opt_done, opt_stream));
      cx.stream = prev;
      cx.sm_budget = 0;
      opt_bwd_pending = true;
    }
  } else {
    VD_REQUIRE(dlogits, VD_E_STATE, "decoder_backward before criterion_backward");
    float* do2 = arena.get<float>(N * db.Ta * H);
    linear_bwd(seg("dec.out.weight"), dec2.h, dlogits, N * db.Ta, do2, 0.f);
    for (int l = 0; l < 2; ++l) { gen_dh0[l] = arena.get<float>(N * H); gen_dc0[l] = arena.get<float>(N * H); }
    float* dx2 = arena.get<float>(N * db.Ta * H);
    lstm_backward(dec2, do2, nullptr, nullptr, dx2, gen_dh0[1], gen_dc0[1]);
    float* dx1 = arena.get<float>(N * db.Ta * E);
    lstm_backward(dec1, dx2, nullptr, nullptr, dx1, gen_dh0[0], gen_dc0[0]);
    embed_scatter_add(cx, dWp(0), dx1, E, ids_ai, N * db.Ta, E, dropcfg(0.f), 0);
  }
}

const float* Engine::backward_connect() {
  // decoders/gen.lua:45-60
  if (cfg.dec == DEC_DISC) return dEncFromDec;     // t[2] of model.lua:335
  conn_dh_l1 = conn_dc_l1 = conn_dc_l2 = nullptr;
  if (cfg.rnn_layers) {
    conn_dc_l1 = gen_dc0[0]; conn_dc_l2 = gen_dc0[1];
    conn_dh_l1 = gen_dh0[0];
  }
  return gen_dh0[1];
}

// Model:retrieveBatch (model.lua:344-430)
void Engine::retrieve(const vd_batch* b, int use_gt, int32_t* ranks_host) {
  VD_REQUIRE(ranks_host != nullptr, VD_E_BADARG, "ranks_host is null");
  int saved_training = training;
  training = 0;
  try {
    encoder_forward(b);
    const int64_t N = db.N;
    const float* sc_dev = nullptr;
    if (cfg.dec == DEC_DISC) { decoder_forward(); sc_dev = scores; }
    else { gen_option_lhood(); sc_dev = lhood; }
    if (use_gt) VD_REQUIRE(db.answer_ind, VD_E_SHAPE, "use_gt needs answer_ind");
    int64_t nout = use_gt ? N : N * cfg.K;
    int32_t* ranks = arena.get<int32_t>(nout);
    rank_rows(cx, sc_dev, use_gt ? db.answer_ind : nullptr, ranks, N, cfg.K);
    VD_CUDA_CHECK(cudaMemcpyAsync(ranks_host, ranks, (size_t)nout * sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
    VD_CUDA_CHECK(cudaStreamSynchronize(cx.stream));
  } catch (...) { training = saved_training; throw; }
  training = saved_training;
}

// gen retrieval: the 100-iteration option loop of model.lua:405-415 batched over (N*100) rows, log-likelihood
// accumulated per time step without materialising (T,N*100,V) log-probs.
void Engine::gen_option_lhood() {
  VD_REQUIRE(cfg.dec == DEC_GEN, VD_E_STATE, "gen_option_lhood needs the gen decoder");
  VD_REQUIRE(have_fwd && db.option_in && db.option_out && db.To > 0, VD_E_SHAPE, "option_in / option_out / To missing");
  const int E = cfg.E, H = cfg.H, K = cfg.K, V = cfg.V;
  const int64_t N = db.N, Ro = N * K;
  forward_connect();
  int32_t* oi = arena.get<int32_t>(Ro * db.To);
  int32_t* oo = arena.get<int32_t>(Ro * db.To);
  transpose_ids(cx, db.option_in, oi, Ro, db.To);
  transpose_ids(cx, db.option_out, oo, Ro, db.To);
  // (h0,c0) repeated over the K options of a round: row (n,k) <- row n
  const float* h0[2] = {nullptr, nullptr}; const float* c0[2] = {nullptr, nullptr};
  for (int l = 0; l < 2; ++l) {
    if (gen_h0[l]) { float* t = arena.get<float>(Ro * H); repeat_rows(cx, t, gen_h0[l], N, K, H); h0[l] = t; }
    if (gen_c0[l]) { float* t = arena.get<float>(Ro * H); repeat_rows(cx, t, gen_c0[l], N, K, H); c0[l] = t; }
  }
  LstmRun l1 = make_run(db.To, Ro, E, H, seg("dec.lstm1.weight"), nullptr, oi, oi);
  l1.h0 = h0[0]; l1.c0 = c0[0];
  lstm_forward(l1, true);
  LstmRun l2 = make_run(db.To, Ro, H, H, seg("dec.lstm2.weight"), l1.h, nullptr, oi);
  l2.h0 = h0[1]; l2.c0 = c0[1];
  lstm_forward(l2, true);
  lhood = arena.get<float>(Ro);
  VD_CUDA_CHECK(cudaMemsetAsync(lhood, 0, (size_t)Ro * sizeof(float), cx.stream));
  int wo = seg("dec.out.weight");
  // utils.computeLhood per time step.  Tensor-core modes: the vocabulary projection keeps its (N*100, V) logits on chip and
  // writes (max, sum exp) partials + the target logit only (20 MB per step instead of 1.28 GB written and read back)
  const int nparts = vocab_lse_nparts(V);
  float *pm = nullptr, *ps = nullptr, *tl = nullptr, *logits = nullptr;
  bool fused = tcmode() && Ro >= 64 && V >= 256;
  if (fused) { pm = arena.get<float>(Ro * nparts); ps = arena.get<float>(Ro * nparts); tl = arena.get<float>(Ro); }
  for (int t = 0; t < db.To; ++t) {
    const float* ht = l2.h + (int64_t)t * Ro * H;
    if (fused) {
      LaunchCtx::Scope sc(&cx, "vocab_lse", 2.0 * Ro * V * H, 4.0 * (Ro * (double)H + (double)V * H + 2.0 * Ro * nparts));
      fused = vocab_lse_tc(cx, (int)Ro, V, H, ht, H, Wp(wo), H, Wp(wo + 1), oo + (int64_t)t * Ro, pm, ps, tl);
      VD_REQUIRE(fused || t == 0, VD_E_STATE, "vocab_lse_tc changed its mind between time steps");
    }
    if (fused) {
      vocab_lse_finish(cx, pm, ps, nparts, tl, oo + (int64_t)t * Ro, oi + (int64_t)t * Ro, nullptr, lhood, 1.f, 1, Ro);
    } else {
      if (!logits) logits = arena.get<float>(Ro * V);
      linear_fwd(wo, ht, Ro, logits, 0);
      lhood_accumulate(cx, logits, oo + (int64_t)t * Ro, oi + (int64_t)t * Ro, lhood, Ro, V);
    }
  }
}

// Model:generateAnswers' inner decoder call (model.lua:517-526): decoders/gen.lua:3-27 for ONE time step on `rows`
// independent rows with explicit previous state.  Same kernels, in the same order, as decoder_forward's gen branch with
// Ta = 1 (dense embedded input, maskzero on the token, MaskZero'd Linear + LogSoftMax).
void Engine::gen_decoder_step(int64_t rows, const int32_t* tokens_host, const float* const* h_prev, const float* const* c_prev) {
  VD_REQUIRE(cfg.dec == DEC_GEN, VD_E_STATE, "gen_decoder_step needs the gen decoder");
  VD_REQUIRE(have_fwd, VD_E_STATE, "gen_decoder_step before encoder_forward");
  VD_REQUIRE(rows > 0 && tokens_host != nullptr, VD_E_BADARG, "rows / tokens");
  VD_CUDA_CHECK(cudaSetDevice(cfg.gpuid));
  cx.stream = main_stream;
  const int E = cfg.E, H = cfg.H, V = cfg.V;
  int32_t* tok = arena.get<int32_t>(rows);
  VD_CUDA_CHECK(cudaMemcpyAsync(tok, tokens_host, (size_t)rows * sizeof(int32_t), cudaMemcpyHostToDevice, cx.stream));
  VD_CUDA_CHECK(cudaStreamSynchronize(cx.stream));               // tokens_host may be a temporary of the caller
  float* xa = arena.get<float>(rows * E);
  embed_rows(cx, xa, Wp(0), tok, rows, E, dropcfg(0.f), 0);
  gstep1 = make_run(1, rows, E, H, seg("dec.lstm1.weight"), xa, nullptr, tok);
  gstep1.h0 = h_prev ? h_prev[0] : nullptr; gstep1.c0 = c_prev ? c_prev[0] : nullptr;
  lstm_forward(gstep1, true);
  gstep2 = make_run(1, rows, H, H, seg("dec.lstm2.weight"), gstep1.h, nullptr, tok);
  gstep2.h0 = h_prev ? h_prev[1] : nullptr; gstep2.c0 = c_prev ? c_prev[1] : nullptr;
  lstm_forward(gstep2, true);
  gstep_logp = arena.get<float>(rows * V);
  linear_fwd(seg("dec.out.weight"), gstep2.h, rows, gstep_logp, 0);
  logsoftmax_rows(cx, gstep_logp, tok, rows, V);                // gen.lua:23-24 (MaskZero)
}

// One beam-search step (model.lua:510-570) for `rows` hypotheses at once (all rounds of a dialog x beamSize).  The state of
// the previous call stays on the device: parent_host[r] >= 0 takes the state hypothesis parent produced, < 0 the state row
// (-1 - parent) was fed (stale beam column).  parent_host == NULL starts a search from the host arrays init_h / init_c.
void Engine::gen_beam_step(int64_t rows, const int32_t* tokens_host, const int32_t* parent_host, const float* const* init_h_host,
                           const float* const* init_c_host, int k, float* topv_host, int32_t* topi_host) {
  VD_REQUIRE(cfg.dec == DEC_GEN && have_fwd, VD_E_STATE, "gen_beam_step needs the gen decoder after encoder_forward");
  VD_REQUIRE(rows > 0 && tokens_host && topv_host && topi_host && k >= 1 && k <= cfg.V, VD_E_BADARG, "rows / tokens / k");
  VD_CUDA_CHECK(cudaSetDevice(cfg.gpuid));
  cx.stream = main_stream;
  const int H = cfg.H;
  float* hp[2]; float* cp[2];
  for (int l = 0; l < 2; ++l) { hp[l] = arena.get<float>(rows * H); cp[l] = arena.get<float>(rows * H); }
  if (!parent_host) {
    VD_REQUIRE(init_h_host && init_c_host, VD_E_BADARG, "first beam step needs the initial state");
    for (int l = 0; l < 2; ++l) {
      VD_CUDA_CHECK(cudaMemcpyAsync(hp[l], init_h_host[l], (size_t)rows * H * sizeof(float), cudaMemcpyHostToDevice, cx.stream));
      VD_CUDA_CHECK(cudaMemcpyAsync(cp[l], init_c_host[l], (size_t)rows * H * sizeof(float), cudaMemcpyHostToDevice, cx.stream));
    }
  } else {
    VD_REQUIRE(beam_rows == rows && beam_in_h[0], VD_E_STATE, "gen_beam_step: no previous step with this many rows");
    int32_t* par = arena.get<int32_t>(rows);
    VD_CUDA_CHECK(cudaMemcpyAsync(par, parent_host, (size_t)rows * sizeof(int32_t), cudaMemcpyHostToDevice, cx.stream));
    const float* out_h[2] = {gstep1.h, gstep2.h};
    const float* out_c[2] = {gstep1.c, gstep2.c};
    for (int l = 0; l < 2; ++l) {
      beam_gather(cx, hp[l], out_h[l], beam_in_h[l], par, rows, H);
      beam_gather(cx, cp[l], out_c[l], beam_in_c[l], par, rows, H);
    }
  }
  const float* hpc[2] = {hp[0], hp[1]};
  const float* cpc[2] = {cp[0], cp[1]};
  gen_decoder_step(rows, tokens_host, hpc, cpc);            // synchronises once for the token upload
  for (int l = 0; l < 2; ++l) { beam_in_h[l] = hp[l]; beam_in_c[l] = cp[l]; }
  beam_rows = rows;
  float* tv = arena.get<float>(rows * k);
  int32_t* ti = arena.get<int32_t>(rows * k);
  topk_rows(cx, gstep_logp, rows, cfg.V, k, tv, ti);
  VD_CUDA_CHECK(cudaMemcpyAsync(topv_host, tv, (size_t)rows * k * sizeof(float), cudaMemcpyDeviceToHost, cx.stream));
  VD_CUDA_CHECK(cudaMemcpyAsync(topi_host, ti, (size_t)rows * k * sizeof(int32_t), cudaMemcpyDeviceToHost, cx.stream));
  VD_CUDA_CHECK(cudaStreamSynchronize(cx.stream));
}

// model.lua:96-99 + optim_updates.lua:62-91
void Engine::clamp_adam_step(float lr) {
  VD_CUDA_CHECK(cudaSetDevice(cfg.gpuid));
  join_options_backward();
  float gscale = 1.f;
  if (world > 1) {
    allreduce_grads();
    // disc: loss is a mean over the local N rows -> average over ranks; gen: a sum -> plain sum
    if (cfg.dec == DEC_DISC) gscale = 1.f / (float)world;
  }
  adam_t += 1;
  const double b1 = 0.9, b2 = 0.999;
  double bc1 = 1.0 - pow(b1, (double)adam_t), bc2 = 1.0 - pow(b2, (double)adam_t);
  float step = (float)((double)lr * sqrt(bc2) / bc1);
  clamp_adam(cx, W, dW, m, v, nparams, step, (float)b1, (float)b2, 1e-8f, gscale);
}

}  // namespace vd
