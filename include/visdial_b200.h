/*
 * visdial_b200 — C ABI of the B200-native Visual Dialog encoder/decoder engine.
 *
 * This is the drop-in boundary for the reference's per-batch hot path.  Every entry point cites
 * the reference interface it replaces (paths relative to /root/reference).  Host languages bind it
 * directly: LuaJIT `ffi.cdef` (see INTEGRATION.md and lua/), Python ctypes (visdial_b200/_lib.py).
 *
 * Conventions
 *  - every call returns int: 0 = ok, <0 = error class (VD_E_*); message via vd_last_error().
 *    No exceptions cross the ABI, nothing calls exit().  (The reference's error()/assert kill the
 *    CLI: model.lua:436, weight-init.lua:46; the Lua shim turns rc != 0 into error(msg).)
 *  - all tensors fp32; token / class ids int32, 1-based with 0 = pad, exactly as the reference
 *    dataloader emits them (dataloader.lua:143-321; on GPU the reference stores them as fp32).
 *  - batch tensors are passed BATCH-MAJOR as the dataloader hands them to Model:forwardBackward
 *    (model.lua:255-311 only re-views them time-major; the engine does that indexing itself).
 *  - one engine <-> one device <-> one non-default stream; an engine is single-caller.
 *  - pointers returned by the engine (parameters, outputs) are DEVICE pointers that stay valid for
 *    the engine's lifetime (outputs: until the next call that produces the same output).
 */
#ifndef VISDIAL_B200_H
#define VISDIAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VD_OK 0
#define VD_E_BADARG (-1)   /* null pointer, unknown encoder/decoder name, bad flag            */
#define VD_E_SHAPE (-2)    /* batch sizes inconsistent with params                             */
#define VD_E_CUDA (-3)     /* CUDA runtime / launch failure                                    */
#define VD_E_COMM (-4)     /* NCCL failure                                                     */
#define VD_E_OOM (-5)      /* device allocation failed                                         */
#define VD_E_STATE (-6)    /* call order violated (backward before forward, ...)               */

typedef struct vd_engine vd_engine;

/* modelParams (opts.lua:6-40, train.lua:55-59, evaluate.lua:72-75).  Field names are the
 * reference's.  useHistory / useIm / concatHistory are derived from `encoder` by substring match
 * exactly as opts.lua:55-59 does. */
typedef struct vd_params {
  const char* encoder;          /* 'lf-ques' | 'lf-ques-im-hist' | 'hrea-ques-im-hist' | 'mn-att-ques-im-hist' */
  const char* decoder;          /* 'disc' | 'gen' */
  int32_t vocabSize;            /* V, incl. <START>=V-1 and <END>=V; the embedding has V+1 rows */
  int32_t embedSize;            /* 300 */
  int32_t rnnHiddenSize;        /* 512 */
  int32_t numLayers;            /* 2 (the only value the configured graphs are built for) */
  int32_t imgFeatureSize;       /* 4096 (fc7) or 512 (pool5 channels) */
  int32_t imgSpatialSize;       /* 14 */
  int32_t imgEmbedSize;         /* 300 */
  int32_t commonEmbeddingSize;  /* 512 */
  int32_t numAttentionLayers;   /* 1 */
  int32_t maxQuesCount;         /* 10 rounds */
  int32_t numOptions;           /* 100 */
  float dropout;                /* 0.5; used by the LF graphs, MN/att/HREA hard-code 0.5 */
  int32_t gpuid;                /* CUDA device ordinal */
} vd_params;

/* A batch as produced by dataloader:getTrainBatch / getTestBatch (dataloader.lua:324-478),
 * Appendix A of SURVEY.md.  Unused pointers may be NULL.  N = B * maxQuesCount. */
typedef struct vd_batch {
  int32_t B;                    /* dialogs ("threads") in this batch */
  int32_t Tq, Th, Ta, To;       /* trimmed question / history / answer(+1) / option widths */
  const int32_t* ques_fwd;      /* (B,10,Tq) right-aligned */
  const int32_t* hist;          /* (B,10,Th) right-aligned */
  const float* img_feat;        /* (B,F) fc7 or (B,S,S,C) NHWC pool5 — one row per dialog */
  const int32_t* options;       /* disc: (N,100,To) left-aligned raw option tokens */
  const int32_t* answer_ind;    /* (N) 1-based ground-truth option */
  const int32_t* answer_in;     /* gen: (B,10,Ta) <START> a.. 0.. */
  const int32_t* answer_out;    /* gen: (B,10,Ta) a.. <END> 0.. */
  const int32_t* option_in;     /* gen eval: (B,10,100,To) */
  const int32_t* option_out;    /* gen eval: (B,10,100,To) */
  int32_t on_device;            /* 0: host pointers (staged + copied H2D by the engine), 1: device */
} vd_batch;

/* ---- parameter layout: host-only, needs no GPU --------------------------------------------- */
/* Replaces nn.Module:getParameters() flattening (model.lua:55).  Segment order = DESIGN.md §3. */
#define VD_INIT_EMBED 0       /* N(0,1)                       [upstream nn.LookupTable]            */
#define VD_INIT_LINEAR_W 1    /* U(-1/sqrt(in), 1/sqrt(in))   [upstream nn.Linear]                 */
#define VD_INIT_LINEAR_B 2    /* U(-1/sqrt(in), 1/sqrt(in))                                        */
#define VD_INIT_LSTM_W 3      /* N(0, 1/sqrt(D+H))            [upstream rnn.SeqLSTM]               */
#define VD_INIT_LSTM_B 4      /* 0, forget block (cols H..2H) = 1                                  */
int vd_layout_count(const vd_params* p, int32_t* n_segments, int64_t* n_params);
int vd_layout_segment(const vd_params* p, int32_t idx, char* name, int32_t name_cap,
                      int64_t* offset, int64_t* rows, int64_t* cols, int32_t* init_kind,
                      int64_t* fan_in);

/* ---- lifetime ------------------------------------------------------------------------------ */
/* encoder.model(params) + decoder.model(params, enc) + nn.Sequential wrapper + :cuda()
 * (model.lua:19-55). */
int vd_create(const vd_params* p, vd_engine** out);
int vd_destroy(vd_engine* e);
const char* vd_last_error(void);

/* ---- parameters (wrapperW / wrapperdW, model.lua:55) --------------------------------------- */
int vd_num_params(vd_engine* e, int64_t* n);
int vd_param_buffers(vd_engine* e, float** W_dev, float** dW_dev);
int vd_optim_buffers(vd_engine* e, float** m_dev, float** v_dev, int64_t* t);
/* restore Adam's state table (optims.m / .v / .t of model_utils/optim_updates.lua:67-84) from HOST vectors of
 * vd_num_params floats — resuming from a checkpoint written by train.lua:99-102 */
int vd_set_optim_state(vd_engine* e, const float* m_host, const float* v_host, int64_t t);
int vd_set_parameters(vd_engine* e, const float* host_src, int64_t n);   /* wrapperW:copy(modelW), evaluate.lua:91 */
int vd_get_parameters(vd_engine* e, float* host_dst, int64_t n);
int vd_get_gradients(vd_engine* e, float* host_dst, int64_t n);
int vd_zero_grad(vd_engine* e);                                          /* wrapper:zeroGradParameters(), model.lua:68 */

/* ---- modes ---------------------------------------------------------------------------------- */
int vd_set_training(vd_engine* e, int32_t training);   /* wrapper:training()/:evaluate(), model.lua:57,111 */
/* Dropout masks are a pure function philox4x32-10(seed; site, iteration, element index)
 *  so that the oracle can be given identical masks. */
int vd_set_dropout_seed(vd_engine* e, uint64_t seed, uint64_t iteration);
#define VD_MATH_TF32 0        /* dense contractions on tcgen05 tensor cores, TF32 operands, fp32 accumulate */
#define VD_MATH_FP32 1        /* same contractions on CUDA cores in fp32 (verification mode) */
#define VD_MATH_F16 2         /* TF32 mode + the many-row option LSTM (disc.lua:4-20) with fp16 operands and fp16 saved state
                                 (h, gates, da, x-projection table), fp32 accumulation, fp32 cell state and gradients */
int vd_set_math_mode(vd_engine* e, int32_t mode);
/* gen decoder: on = 1 lets vd_decoder_forward keep the (rows, vocabSize) log-probabilities on chip when the caller only
 * passes decOut on to the criterion (decoder:forward -> criterion:forward, model.lua:313-314): decOut_dev is then NULL, the
 * projection's epilogue keeps the softmax statistics and the target log-probability, and vd_criterion_backward recomputes the
 * projection with the softmax gradient fused in.  Default 0: decOut is materialised (LogSoftMax output, gen.lua:23-24).
 * vd_forward_backward and vd_retrieve never hand decOut out and always take the fused route in the tensor-core modes. */
int vd_set_lazy_decout(vd_engine* e, int32_t on);
/* Scheduling knob (results are identical either way).  on = 1 (default): the disc decoder's option LSTM (disc.lua:4-20),
 * which does not depend on the encoder until the final dot product, runs on its own stream concurrently with the
 * encoder's forward and backward, its persistent kernels leaving `reserve_sms` SMs (default 16, < 0 keeps the current
 * value) to the encoder's chains.  on = 0: the reference's order (encoder, then decoder) on one timeline — used by
 * bench.py to time the option-LSTM step kernel alone for the roofline. */
int vd_set_option_overlap(vd_engine* e, int32_t on, int32_t reserve_sms);

/* ---- fine-grained module protocol (what Model:forwardBackward calls, model.lua:297-337) ----- */
int vd_encoder_forward(vd_engine* e, const vd_batch* b, const float** encOut_dev);       /* encoder:forward(inputs), :297 */
int vd_forward_connect(vd_engine* e);                                                    /* decoders/gen.lua:30-42; no-op for disc (disc.lua:35) */
int vd_decoder_forward(vd_engine* e, const vd_batch* b, const float** decOut_dev);       /* decoder:forward, :313 / :329 */
int vd_criterion_forward(vd_engine* e, const vd_batch* b, float* loss_host);             /* criterion:forward, :314 / :330 */
int vd_criterion_backward(vd_engine* e, const vd_batch* b);                              /* criterion:backward, :318 / :334 */
int vd_decoder_backward(vd_engine* e, const vd_batch* b);                                /* decoder:backward, :319 / :335 */
int vd_backward_connect(vd_engine* e, const float** gradEncOut_dev);                     /* gen.lua:45-60; disc: t[2] of :335 */
int vd_encoder_backward(vd_engine* e, const vd_batch* b, const float* gradEncOut_dev);   /* encoder:backward, :323 / :337 */

/* ---- fused fast paths ----------------------------------------------------------------------- */
int vd_forward_backward(vd_engine* e, const vd_batch* b, int32_t only_forward, float* loss_host);  /* Model:forwardBackward, model.lua:249-342 */
/* Model:retrieveBatch (model.lua:344-430).  use_gt != 0: ranks_host is (N) = rank of the ground
 * truth; else (N,100) = rank of every option.  Ranks are 1-based; ties: lower index wins. */
int vd_retrieve(vd_engine* e, const vd_batch* b, int32_t use_gt, int32_t* ranks_host);
/* utils.computeRanks (utils.lua:106-128) on device scores (n_rows,100). */
int vd_compute_ranks(vd_engine* e, const float* scores_dev, int32_t n_rows,
                     const int32_t* gt_dev_or_null, int32_t* ranks_dev);
/* utils.computeLhood (utils.lua:86-102) fused with the gen decoder: log-likelihood (N,100) of
 * every candidate answer, never materialising the (T,N,V) log-probs. */
int vd_gen_option_lhood(vd_engine* e, const vd_batch* b, const float** lhood_dev);
/* Model:generateAnswers (model.lua:432-613): the pieces its beam search / sampling loop drives on the device.
 *  vd_encoder_rnn_state: enc.rnnLayers[level].output[Tq] / .cell[Tq] of the last vd_encoder_forward, (N,H) device
 *    pointers, level = 0 | 1 (model.lua:480-483); both NULL for encoders without .rnnLayers (mn-att, :491-501).
 *  vd_gen_decoder_step: decoder:forward(tokens) for ONE time step on `rows` independent rows with
 *    .userPrevOutput / .userPrevCell = h_prev[l] / c_prev[l] (l = 0, 1; (rows,H) device pointers, NULL = zeros)
 *    (model.lua:517-526, gen.lua:3-27).  tokens: HOST int32 (rows).  Returns device pointers, valid until the next
 *    vd_encoder_forward: log-probabilities (rows,V) (all-zero row for a pad token, MaskZero) and the new state. */
int vd_encoder_rnn_state(vd_engine* e, int32_t level, const float** h_last_dev, const float** c_last_dev);
int vd_gen_decoder_step(vd_engine* e, int32_t rows, const int32_t* tokens_host, const float* const* h_prev,
                        const float* const* c_prev, const float** logp_dev, const float** h_out, const float** c_out);
/* Beam search with the search state on the device (model.lua:510-570, all rounds of a dialog at once): one decoder step on
 * `rows` hypotheses.  parent_host == NULL starts a search: init_h_host / init_c_host[2] are (rows, rnnHiddenSize) HOST arrays
 * (model.lua:480-501).  Otherwise parent_host[r] >= 0 continues hypothesis r from the state row `parent` PRODUCED in the
 * previous call, parent_host[r] < 0 from the state row (-1 - parent) was FED in the previous call (a beam column that received
 * no candidate keeps its old content, :560-569).  Only the k best (log-prob, 0-based class) pairs of every row come back
 * (torch.topk sorted; ties: lower class first) — the (rows, vocabSize) log-probabilities and the LSTM state stay in HBM. */
int vd_gen_beam_step(vd_engine* e, int32_t rows, const int32_t* tokens_host, const int32_t* parent_host,
                     const float* const* init_h_host, const float* const* init_c_host, int32_t k, float* topv_host,
                     int32_t* topi_host);

/* ---- optimiser step (model.lua:96-105 + optim_updates.lua:62-91) ---------------------------- */
/* all-reduce(SUM)/world of dW when a communicator is attached, then clamp(-5,5), then adam.
 * The LR decay (model.lua:102-105) stays with the caller, as in the reference. */
int vd_clamp_adam_step(vd_engine* e, float learning_rate);

/* ---- data-parallel communicator (no reference counterpart: train.lua is single-GPU) -------- */
#define VD_COMM_ID_BYTES 128
int vd_comm_unique_id(void* id_out);                        /* rank 0; broadcast the bytes out of band */
int vd_comm_init(vd_engine* e, const void* id, int32_t rank, int32_t world);
int vd_comm_allreduce_grads(vd_engine* e);                  /* exposed for tests; vd_clamp_adam_step calls it */

/* ---- plumbing -------------------------------------------------------------------------------- */
int vd_memcpy_d2h(vd_engine* e, void* host_dst, const void* dev_src, size_t bytes);
int vd_memcpy_h2d(vd_engine* e, void* dev_dst, const void* host_src, size_t bytes);
/* pinned host memory for batch buffers (the H2D copy of a batch is only asynchronous from pinned memory) */
int vd_host_alloc(void** ptr, size_t bytes);
int vd_host_free(void* ptr);
/* device memory for callers that keep batches resident in HBM (vd_batch.on_device = 1) */
int vd_device_alloc(vd_engine* e, void** ptr, size_t bytes);
int vd_device_free(vd_engine* e, void* ptr);
int vd_synchronize(vd_engine* e);
int vd_stream(vd_engine* e, void** cuda_stream);
/* device-side timing on the engine's stream (CUDA events) */
int vd_timer_start(vd_engine* e);
int vd_timer_stop(vd_engine* e, float* ms);
/* launch accounting: every kernel the engine launches is counted; kernels of class `name`
 * ("lstm_step", "gemm", ...) are additionally bracketed by events when profiling is on. */
int vd_profile_enable(vd_engine* e, int32_t on);
int vd_profile_reset(vd_engine* e);
int vd_launch_count(vd_engine* e, int64_t* n_launches);
int vd_kernel_stats(vd_engine* e, const char* name, int64_t* launches, double* total_ms,
                    double* total_flops, double* total_bytes);
/* test hooks: the engine's two dense-contraction primitives on caller-provided DEVICE buffers, routed exactly as
 * the engine routes them (math mode).  tn: C[m,n] = act(beta*C + bias[n] + sum_k A[m,k] B[n,k]);
 * atb: C[m,n] += sum_k A[k,m] B[k,n]. */
int vd_gemm_tn(vd_engine* e, int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
               float* C, int64_t ldc, float beta, const float* bias, int32_t act);
int vd_gemm_atb(vd_engine* e, int32_t M, int32_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                float* C, int64_t ldc);
/* test hook of the VD_MATH_F16 weight-gradient primitive: A (K x M) and B (K x N) fp32 DEVICE buffers are rounded to
 * fp16, then C[m,n] += inv_scale * sum_k A[k,m] B[k,n] on tcgen05 kind::f16 (both operands MN-major). */
int vd_gemm_atb16(vd_engine* e, int32_t M, int32_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                  float* C, int64_t ldc, float inv_scale);
/* cudaProfilerStart / cudaProfilerStop (ncu --profile-from-start off) */
int vd_profiler_range(vd_engine* e, int32_t start);
/* flush L2 by writing a scratch buffer larger than L2 (bench hygiene) */
int vd_flush_l2(vd_engine* e);

/* ---- dataloader: HBM-resident corpus and on-device batch assembly --------------------------------
 * Replaces dataloader:initialize's tensor preparation and the per-batch indexing
 * (dataloader.lua:143-321 prepareDataset / processAnswers / processHistory / processOptions,
 * utils.lua:6-45 rightAlign, dataloader.lua:324-478 getTrainBatch / getTestBatch / getIndexData /
 * getIndexOption).  The raw arrays are the datasets of visdial_data.h5 for ONE split exactly as
 * data/prepro.py:105-183 writes them (the HDF5 read itself stays with the host language); they are
 * uploaded once, prepared on the device, and every later batch is gathered + trimmed on the device:
 * per batch only the dialog indices cross PCIe.  All ids int32, 1-based with 0 = pad. */
typedef struct vd_corpus vd_corpus;

typedef struct vd_corpus_desc {
  int32_t numThreads;           /* dialogs in the split: ques:size(1)                      (dataloader.lua:94-105) */
  int32_t numRounds;            /* ques:size(2) = maxQuesCount                             (:122)  */
  int32_t maxQuesLen;           /* ques:size(3)                                            (:124)  */
  int32_t maxAnsLen;            /* ans:size(3) = opt_list:size(2)                          (:126)  */
  int32_t maxCapLen;            /* cap:size(2); must be >= maxQuesLen + maxAnsLen when useHistory (:236) */
  int32_t numOptions;           /* opt:size(3)                                             (:112)  */
  int32_t numOptList;           /* opt_list:size(1)                                                */
  int32_t numImages;            /* images:size(1)                                                  */
  int32_t useHistory;           /* opt.useHistory / concatHistory / useIm                  (:139-141) */
  int32_t concatHistory;
  int32_t useIm;
  int32_t maxHistoryLen;        /* opt.maxHistoryLen (60) — overwritten by min(R*(Lq+La),300) when concatHistory (:142,:217) */
  int32_t imgNorm;              /* 1: L2-normalise over dim 2 at load                      (:64-68) */
  int32_t imgAtt;               /* 1: images are (N,C,S,S) and are stored (N,S,S,C)        (:70-72) */
  int32_t imgChannels;          /* C (pool5 512) or F (fc7 4096)                                   */
  int32_t imgSpatial;           /* S (14); ignored unless imgAtt                                   */
  int32_t startToken, endToken; /* word2ind['<START>'], word2ind['<END>']                  (:17-22) */
  const int32_t* ques;          /* (n,R,Lq) left-aligned  'ques_<split>'                           */
  const int32_t* ques_len;      /* (n,R)                  'ques_length_<split>'                    */
  const int32_t* ans;           /* (n,R,La)               'ans_<split>'                            */
  const int32_t* ans_len;       /* (n,R)                  'ans_length_<split>'                     */
  const int32_t* cap;           /* (n,Lc)                 'cap_<split>'        (NULL unless useHistory) */
  const int32_t* cap_len;       /* (n)                    'cap_length_<split>'                     */
  const int32_t* opt;           /* (n,R,K) 1-based rows of opt_list  'opt_<split>'                 */
  const int32_t* opt_list;      /* (m,La)                 'opt_list_<split>'                       */
  const int32_t* opt_len;       /* (m)                    'opt_length_<split>'                     */
  const int32_t* ans_index;     /* (n,R) 1-based gt option 'ans_index_<split>' (NULL for test)     */
  const int32_t* img_pos;       /* (n) 0-based row of images as stored ('img_pos_<split>'; the +1 of :76-77 is Lua indexing) */
  const int32_t* num_rounds;    /* (n)                    'num_rounds_<split>' (may be NULL)       */
  const float* images;          /* (N,F) or (N,C,S,S) fp32 'images_<split>' of the image h5 (NULL unless useIm) */
} vd_corpus_desc;

/* dataloader:initialize tensors -> HBM + dataloader:prepareDataset on the device.  Host pointers. */
int vd_corpus_create(vd_engine* e, const vd_corpus_desc* d, vd_corpus** out);
int vd_corpus_destroy(vd_corpus* c);
/* Assemble the batch of dialogs `inds` (HOST array of n 0-based dialog indices = Lua's inds - 1) on the engine's stream.
 *   decoder_gen = 0: getIndexData + getIndexOption('disc')  -> ques_fwd, hist, img_feat, answer_in/out, answer_ind, options
 *   decoder_gen = 1: getIndexData only (getTrainBatch, gen)  -> ... without options
 *   decoder_gen = 2: getIndexData + getIndexOption('gen')   -> ... plus option_in / option_out (getTestBatch, gen)
 * `out` receives DEVICE pointers (on_device = 1) into corpus-owned buffers that stay valid until the second-next
 * vd_corpus_get_batch on this corpus (two buffer sets, alternating). */
int vd_corpus_get_batch(vd_corpus* c, const int64_t* inds, int32_t n, int32_t decoder_gen, vd_batch* out);
/* Read one PREPARED tensor back to the host (parity tests): name in {"ques_fwd","hist","hist_len","ans_in","ans_out",
 * "opt_in","opt_out","img_fv"}; *elems receives the element count (call with host_dst = NULL to size). */
int vd_corpus_read(vd_corpus* c, const char* name, void* host_dst, int64_t* elems);
/* bytes the last vd_corpus_get_batch read + wrote in HBM (algorithmic) and the number of kernels it launched */
int vd_corpus_batch_bytes(vd_corpus* c, int64_t* bytes, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* VISDIAL_B200_H */
