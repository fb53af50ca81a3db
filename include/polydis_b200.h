/* libpolydis_b200 -- C ABI of the PolyDis hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (ZZWaang/polyphonic-chord-texture-disentanglement) is pure Python on PyTorch ATen; it has
 * no FFI of its own.  The entry points below are the native boundary a binding for its hot path needs:
 * each one replaces an ATen call site of the reference (cited file:line), takes plain device pointers,
 * element strides and sizes, launches on the given CUDA stream and returns immediately.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (fp32 unless noted); strides / leading dimensions are in ELEMENTS
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises or allocates device memory
 *   - return value: 0 = launched; > 0 = cudaError_t; PD_BAD_ARG (-22) = argument contract violated
 *     (for pd_gemm_tf32: operands TMA cannot address -- call pd_gemm_f32 instead)
 *   - buffers documented "accumulated" must be initialised by the caller
 *   - thread safety: entry points hold no mutable global state besides one-time function attributes;
 *     concurrent calls on different streams are safe
 */
#ifndef POLYDIS_B200_H
#define POLYDIS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PD_BAD_ARG (-22)

/* ---- GEMM: nn.Linear / GRU projections (reference: every nn.Linear and the matmuls inside aten::gru,
 * ptvae.py:26-27,63-68,115-120,343,349-352,361,374-375,396-398,435-437,461-462).
 *   C[m*ldc+n] (+)= sum_k A(m,k) B(k,n) (+ bias[n]);  A(m,k)=A[m*sam+k*sak], B(k,n)=B[k*sbk+n*sbn];
 *   one of (sam,sak) and one of (sbk,sbn) must be 1.  accumulate != 0 adds into C (L2 reductions).
 * pd_gemm_f32 : fp32 FFMA, any shape/stride.   pd_gemm_tf32: tcgen05 tensor cores (TF32 x TF32 -> fp32),
 * needs 16-byte aligned bases and row strides that are multiples of 4.  Large plain-store GEMMs leave through TMA bulk
 * stores, which clip at 16-byte granularity: when N % 4 != 0 (and ldc >= N rounded up to 4) the padding columns
 * C[m][N .. roundup4(N)) may be overwritten with zeros. */
int pd_gemm_f32(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                const float* bias, int M, int N, int K, int accumulate, void* stream);
int pd_gemm_tf32(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                 const float* bias, int M, int N, int K, int accumulate, void* stream);
/* tuning variant of pd_gemm_tf32: cfg = BN*100 + stages*10 + CTAs/SM of an instantiated tile config (0 = heuristic) */
int pd_gemm_tf32_cfg(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                     const float* bias, int M, int N, int K, int accumulate, int cfg, void* stream);
/* bf16 operands (bit patterns; strides in elements, multiples of 8), fp32 accumulate / output: kind::f16 MMAs */
int pd_gemm_bf16(const void* A, long sam, long sak, const void* B, long sbk, long sbn, float* C, long ldc,
                 const float* bias, int M, int N, int K, int accumulate, void* stream);
int pd_f32_to_bf16(const float* x, long ldx, long rows, int cols, void* out, long ldo, void* stream);
/* hi = round_to_nearest_tf32(x), lo = x - hi (row stride ldo): operands of the error-compensated "3xTF32" GEMM
 * A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi (TF32 multiplies, fp32 accumulation: ~fp32 accuracy on tensor cores) */
int pd_tf32_split(const float* x, long ldx, long rows, int cols, float* hi, float* lo, long ldo, void* stream);
/* single-launch form: out (rows, 3*kp), kp = cols rounded up to 4: [hi | hi | lo] (order 0, A side) or [hi | lo | hi]
 * (order 1, B side); one pd_gemm_tf32 over K = 3*kp then equals the three-product sum */
int pd_tf32_split3(const float* x, long ldx, long rows, int cols, float* out, long ldo, int order, void* stream);
/* out[n] (+)= sum_m X[m*ldx+n]   (bias gradients) */
int pd_colsum_f32(const float* X, long ldx, int M, int N, float* out, int accumulate, void* stream);
/* the same sum over the live rows only of an (R, T, N) buffer of a length-masked recurrence: row (r, t) is read iff
 * t < lengths[r] (the dead rows hold zeros).  16-byte aligned X, ldx % 4 == 0, N % 4 == 0. */
int pd_colsum_seq_f32(const float* X, long ldx, int R, int T, int N, const int* lengths, float* out, int accumulate,
                      void* stream);
int pd_transpose_f32(const float* in, int rows, int cols, float* out, void* stream);
/* out[r*ldo+c] = sum_t X[r*ldr + t*ldt + c]: gradient of a projection broadcast over a GRU's steps (the hoisted
 * z / summary terms of ptvae.py:394-395,458-460) in one pass instead of a read-modify-write per step */
int pd_sum_steps_f32(const float* X, long ldr, long ldt, int T, float* out, long ldo, long R, int C, void* stream);

/* ---- GRU cell gate math around the per-step W_hh GEMM (aten::gru at ptvae.py:63-65,359-360,396-398,
 * 461-462; packed bi-GRU at :446-453,:480-486 through the per-row `lengths` mask).  Gate order r|z|n.
 *   gi (B,3H) x-projection incl. b_ih; gi2 optional second term (same shape, e.g. the sequence-constant
 *   half of the reference's torch.cat input); gh (B,3H) = W_hh h + b_hh; hprev NULL = zeros.
 *   rzn/hn (may be NULL) receive r,z,n and (W_hn h + b_hn) for the backward pass.
 *   Rows with t >= lengths[b] carry their state (lengths NULL = no mask). */
int pd_gru_gates_fwd(const float* gi, long ldgi, const float* gi2, long ldgi2, const float* gh, long ldgh,
                     const float* hprev, long ldhp, float* hout, long ldho, float* rzn, long ldrzn, float* hn,
                     long ldhn, const int* lengths, int t, int B, int H, void* stream);
/* inference variant: also writes h3 (B, 3H; row stride ldh3) = [hi | hi | lo] TF32 split of the new state, i.e. the
 * pd_tf32_split3 (order 0) operand of the 3xTF32 GEMMs that consume the state next; no backward saves */
int pd_gru_gates_fwd_split3(const float* gi, long ldgi, const float* gi2, long ldgi2, const float* gh, long ldgh,
                            const float* hprev, long ldhp, float* hout, long ldho, const int* lengths, int t, int B, int H,
                            float* h3, long ldh3, void* stream);
/* d = dh + dh2 + dh3 (each optional): dgi=[dr,dz,dn], dgh=[dr,dz,dn*r], dhprev = d*z (the caller's GEMM adds
 * dgh W_hh), dgi2 (optional) += dgi. */
int pd_gru_gates_bwd(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3, long lddh3,
                     const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp,
                     float* dgi, long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, float* dgi2,
                     long lddgi2, const int* lengths, int t, int B, int H, void* stream);

/* fused GRU step on the tensor cores: hout = GRUCell(gi [+ gi2], hprev) with W_hh h computed by tcgen05.mma into
 * TMEM and the gate math done in the epilogue (the (B,3H) h-projection never reaches HBM).  H % 64 == 0,
 * hprev != NULL, hout != hprev, strides multiples of 4 floats, 16-byte aligned bases; else PD_BAD_ARG. */
int pd_gru_step_tf32(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* b_hh, const float* gi,
                     long ldgi, const float* gi2, long ldgi2, float* hout, long ldho, float* rzn, long ldrzn, float* hn,
                     long ldhn, const int* lengths, int t, int B, int H, void* stream);
/* persistent variant of the fused step (two TMEM accumulators, every epilogue operand / result moved as a 32x16 TMA
 * box, operand sets double-buffered so the next chunk's loads fly while a chunk is computed).  Same contract without the
 * length mask; hout must not alias hprev.  Validated on B200 in round 2 (tests/test_gpu_kernels.py::test_gru_step_tma);
 * routed to for recurrences of >= 4096 rows (the 32*B-row note GRU): 108 us vs 147 us for GEMM + gate kernel at
 * 16384 x 512.  pd_gru_step_tma_variant(1) selects the round-1 single-buffered layout (A/B measurements). */
int pd_gru_step_tma(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* b_hh, const float* gi,
                    long ldgi, const float* gi2, long ldgi2, float* hout, long ldho, float* rzn, long ldrzn, float* hn,
                    long ldhn, int B, int H, void* stream);
/* training form with the x-projection folded into the main loop as a second K segment (TF32, single pass): gi = W_x x is never
 * materialised; x (B,K2) rows of the step's input (row stride ldx), w_x (3H,K2) the matching W_ih columns, gi2 (B,3H) the rest
 * of the input projection incl. b_ih.  Teacher-forced note GRU: x = embedding of ground-truth slot n, ptvae.py:396-398. */
int pd_gru_step_tmax(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* x, long ldx, const float* w_x,
                     long ldwx, int K2, const float* b_hh, const float* gi2, long ldgi2, float* hout, long ldho, float* rzn,
                     long ldrzn, float* hn, long ldhn, int B, int H, void* stream);
int pd_gru_step_tma_variant(int variant);
/* inference form of the fused step for the error-compensated 3xTF32 path (greedy decode of >= 512 segments,
 * ptvae.py:396-398,:461-462 at inference): a3 (B,3H) = [hi | hi | lo] split of h_prev, w3 (3H,3H) = [hi | lo | hi] split
 * of W_hh (pd_tf32_split3 orders 0 / 1) -- the three TF32 products accumulate in TMEM; gate math with expf / tanhf; the
 * epilogue writes the new state hout (may alias hprev) and its [hi | hi | lo] split h3out (must not alias a3).  Replaces
 * pd_gemm_tf32 over K = 3H + pd_gru_gates_fwd_split3 (the (B,3H) h-projection never reaches HBM). */
int pd_gru_step_tma3(const float* a3, long lda3, const float* w3, long ldw3, const float* b_hh, const float* gi, long ldgi,
                     const float* gi2, long ldgi2, const float* hprev, long ldhp, float* hout, long ldho, float* h3out,
                     long ldh3, int B, int H, void* stream);
/* the same step with the x-projection folded in as a second K segment of the tcgen05 main loop: x3 (B,K2) = [hi | hi | lo]
 * split of the step's input rows, wx3 (3H,K2) = [hi | lo | hi] split of the matching W_ih columns; gi2 (B,3H) carries the
 * rest of the input projection INCLUDING b_ih.  No (B,3H) x-projection is written or read (note-GRU slot of the greedy
 * decode: x = embedding of the previous token, ptvae.py:396-398,:421-422). */
int pd_gru_step_tma3x(const float* a3, long lda3, const float* w3, long ldw3, const float* x3, long ldx3, const float* wx3,
                      long ldwx3, int K2, const float* b_hh, const float* gi2, long ldgi2, const float* hprev, long ldhp,
                      float* hout, long ldho, float* h3out, long ldh3, int B, int H, void* stream);

/* Whole greedy PianoTree decode (ptvae.py:430-491 with inference=True: 32 time steps x 15 note slots x 5 duration
 * steps, argmax feedback) of B <= 16 segments in ONE cooperative persistent launch: one CTA per SM stays resident, the
 * dependency chain advances through grid barriers in L2, note-GRU / head / summary-GRU weights stay in shared memory.
 * fp32 FFMA arithmetic (token parity with the fp32 reference).  h_time0 (B,1024) = z2dec_hid(z); gi_z (B,3072) =
 * time-GRU projection of z_in incl. b_ih; weights are the state-dict tensors (wt_tok = dec_time_gru.weight_ih[:, :256]
 * with row stride ld_wt, wn_sum / wn_tok = dec_notes_gru.weight_ih[:, :1024] / [:, 1024:] with row stride ld_wn);
 * w_heads (194,512) = [pitch_out_linear | folded dur_hid_linear]; emb_wt (135,128) = note_embedding.weight^T;
 * we_* / be_* = dec_notes_emb_gru forward / reverse.  tokens (32,15,B,6) int32; lens_out (32,B) int32 or NULL;
 * ws: PD_GREEDY_SMALL_WS_FLOATS floats of scratch; bar: 2 x uint32 scratch (bar[1] != 0 afterwards = a grid barrier
 * timed out).  PD_BAD_ARG for B > 16 or when the device cannot co-schedule one CTA per SM. */
#define PD_GREEDY_SMALL_WS_FLOATS 310368
int pd_greedy_decode_small(int B, const float* h_time0, const float* gi_z, const float* wt_tok, long ld_wt,
                           const float* wt_hh, const float* bt_hh, const float* init_tok, const float* w_t2n,
                           const float* b_t2n, const float* wn_sum, long ld_wn, const float* bn_ih, const float* wn_tok,
                           const float* wn_hh, const float* bn_hh, const float* w_heads, const float* b_heads,
                           const float* d_wih, const float* d_bih, const float* d_whh, const float* d_bhh,
                           const float* d_sos, const float* d_wout, const float* d_bout, const float* emb_wt,
                           const float* emb_b, const float* we_ih_f, const float* we_hh_f, const float* be_ih_f,
                           const float* be_hh_f, const float* we_ih_b, const float* we_hh_b, const float* be_ih_b,
                           const float* be_hh_b, int* tokens, int* lens_out, float* ws, unsigned* bar, void* stream);

/* weight-resident variable-length GRU, hidden 128 (the note-summary bi-GRU, ptvae.py:446-453,:480-486): one kernel
 * runs the whole recurrence of a tile of sequences with W_hh resident in shared memory and stops at the tile's
 * longest sequence.  gi (R,T,384) incl. b_ih, lengths (R), h0 = 0; rows past their length carry their state.
 * passes 1 = TF32, 3 = error-compensated TF32.  bwd: dgi / dgh (R,T,384) from dout (R,T,128) and the saves. */
int pd_gru128_fwd(const float* gi, long ldr, long ldt, const int* lengths, const float* w_hh, const float* b_hh,
                  float* h_all, long hr, long ht, float* rzn, long zr, long zt, float* hn, long nr, long nt, long R, int T,
                  int reverse, int passes, void* stream);
/* the same forward with the rows VISITED in the order perm (R) gives (pd_pack_order's: longest first) -- a 16-row tile runs to
 * its longest sequence; every array stays indexed by the original row */
int pd_gru128_fwd_perm(const float* gi, long ldr, long ldt, const int* lengths, const float* w_hh, const float* b_hh,
                       float* h_all, long hr, long ht, float* rzn, long zr, long zt, float* hn, long nr, long nt, long R, int T,
                       int reverse, int passes, const int* perm, void* stream);
int pd_gru128_bwd(const float* dout, long dr, long dt, const float* h_all, long hr, long ht, const float* rzn, long zr,
                  long zt, const float* hn, long nr, long nt, const int* lengths, const float* w_hh, float* dgi, long gr,
                  long gt, float* dgh, long qr, long qt, long R, int T, int reverse, void* stream);

/* ---- PianoTree grid (ptvae.py:292-313,:498-511,:531-535).  x (n_steps,16,6) int64 -> tok int32 (same
 * layout), lengths (n_steps) = 16 - #PAD, pitch targets (n_steps,15), duration targets (n_steps,15,5). */
int pd_grid_prepare(const long long* x, long n_steps, int* tok, int* lengths, int* pitch_tgt, int* dur_tgt,
                    void* stream);
/* data formats either side of the path: pr_mat (n_steps,128) -> grid x (n_steps,16,6) int64
 * (converter.py:116-147 as called in dataset.py:98-104; *overflow set to 1 if a step holds > 14 onsets), and
 * decoded tokens (n_steps,15,6) int32 -> pr_mat (n_steps,128) (ptvae.py:558-575). */
int pd_prmat_to_grid(const float* pr_mat, long n_steps, long long* x, int* overflow, void* stream);
int pd_grid_to_prmat(const int* tok, long n_steps, float* pr_mat, void* stream);
/* decoded tokens int32 (R,6) -> compact uint8 (R,2): [pitch 0..129, the 5 duration bits packed MSB-first]; the
 * device->host format of a decode (2 bytes per note; est_x of ptvae.py:537-544 is rebuilt on the host). */
int pd_pack_tokens(const int* tok, long R, unsigned char* out, void* stream);
/* batch augmentation (dataset.py:67-120): transposition of a segment by shift[b] semitones.  pd_roll_prmat = np.roll
 * of pr_mat (B,32,128) along the pitch axis (converter.py:65-68 augment_pr; out must not alias in);
 * pd_expand_chord = converter.py:150-164 expand_chord: chord rows (rows,14) [root, 12 chroma, bass] -> (rows,36)
 * [root one-hot | rolled chroma | bass one-hot]; rows_per_seg consecutive rows share one shift entry (8 per segment). */
int pd_roll_prmat(const float* pr_in, const int* shift, long B, float* pr_out, void* stream);
int pd_expand_chord(const float* chord14, const int* shift, long rows, int rows_per_seg, float* c36, void* stream);
/* model.py:218-242 interp_path for B latent pairs at once: out (B,count,D); spherical interpolation of the direction,
 * log-linear interpolation of the norm (float64 arithmetic inside, like the reference's numpy). */
int pd_slerp_path(const float* z1, const float* z2, int B, int D, int count, float* out, void* stream);
/* note_embedding(multi-hot) as a gather: out[r] = bias + WT[pitch] (pitch < 130) + sum_k dur_k WT[130+k];
 * tok int32 (R,6); WT = weight^T (135,128).   bwd accumulates dWT (135,128) and dbias (128). */
int pd_note_embed_fwd(const int* tok, long R, const float* WT, const float* bias, float* out, long ldo,
                      void* stream);
int pd_note_embed_bwd(const int* tok, long R, const float* g, long ldg, float* dWT, float* dbias, void* stream);
/* greedy pick for note slot n (ptvae.py:408-416,:425): argmax pitch (first max) and duration bits -> tok row,
 * lens[r] = first n with EOS (15 if none; lens must start at 0). */
int pd_greedy_pick(const float* pitch, long ldp, const float* dur, long ldd, long R, int n, int* tok, long ldtok,
                   int* lens, void* stream);
/* the same pick followed by the embedding of the picked tokens (pd_note_embed_fwd: emb (R,128; row stride lde)) in one
 * launch -- the tail of a greedy note slot (ptvae.py:408-416 + :333) */
int pd_greedy_pick_embed(const float* pitch, long ldp, const float* dur, long ldd, long R, int n, int* tok, long ldtok,
                         int* lens, const float* WT, const float* bias, float* emb, long lde, void* stream);
/* duration feedback token (ptvae.py:322-326): 5-wide, 1 at index == argmax bit */
int pd_dur_token(const float* logit, long ldl, long R, float* tok, void* stream);
/* fused duration decoder (ptvae.py:345-367): 5-step GRU(5->64) + Linear(64->2) with greedy bit feedback.
 * logits (Q,5,2); S (Q,6,72) state buffer for the backward (NULL at inference); GX (Q,6,264) gradient buffer
 * such that GX^T . S holds all parameter gradients (layout: csrc/dur_decoder.cu).  tf32: 0 = fp32 FFMA kernels;
 * 1 = recurrent matvecs on the tensor cores (TF32, training); 3 = error-compensated 3xTF32 matvecs with expf / tanhf
 * gates (fp32-class, the token-parity decode; forward only -- the backward treats any nonzero value as 1). */
int pd_dur_decode_fwd(const float* h0, long ldh0, long Q, const float* w_ih, const float* b_ih, const float* w_hh,
                      const float* b_hh, const float* sos, const float* w_out, const float* b_out, float* logits,
                      float* S, int tf32, void* stream);
/* TF32 calls without saves (S == NULL) of at most n notes run the four-warps-per-tile inference kernel (default 2048;
 * 0 = never): library-wide tuning switch */
int pd_dur_quad_max_notes(int n);
int pd_dur_decode_bwd(const float* S, const float* dlogits, long Q, const float* w_ih, const float* b_ih,
                      const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                      const float* b_out, float* GX, float* dh0, long lddh0, int tf32, void* stream);

/* ---- chord decoder (ptvae.py:73-78; model.py:70-74).  Feedback token = [batch-UNION root one-hot |
 * per-sample chroma argmax | batch-UNION bass one-hot] (the reference's indexing semantics). */
int pd_chord_feedback(const float* root, long ldr, const float* chroma, long ldc, const float* bass, long ldb,
                      int B, float* flags24, float* tok, long ldt, void* stream);
int pd_chord_targets(const float* c, int rows, int* root, int* chroma, int* bass, void* stream);

/* ---- texture encoder front end (ptvae.py:95-99,:114): Conv2d(1->C,(4,12),stride(4,1)) + ReLU +
 * MaxPool(1,4) fused; out (B,C,8,29) channel-major (callers reinterpret as (B,8,29C) like the reference).
 * bwd accumulates dw (C*48) and dbias (C). */
int pd_texture_frontend_fwd(const float* pr_mat, const float* w, const float* bias, int B, int C, float* out,
                            void* stream);
int pd_texture_frontend_bwd(const float* pr_mat, const float* w, const float* bias, int B, int C,
                            const float* gout, float* dw, float* dbias, void* stream);

/* ---- losses (ptvae.py:498-511, model.py:70-90, train_utils.py:33-49).
 * masked mean CE: targets int32, rows with target == ignore skipped; acc2 = {sum, count} kept for bwd. */
int pd_ce_fwd(const float* logits, long ldl, const int* targets, long R, int C, int ignore, float* acc2,
              float* loss, void* stream);
int pd_ce_bwd(const float* logits, long ldl, const int* targets, long R, int C, int ignore, const float* acc2,
              const float* gout, float* dlogits, long lddl, void* stream);
int pd_exp_fwd(const float* x, long n, float* y, void* stream);
int pd_mul_f32(const float* a, const float* b, long n, float* out, void* stream);
int pd_add_f32(const float* a, const float* b, long n, float* out, void* stream);
/* scheduled sampling with a device-resident teacher-forcing plan (ptvae.py:420-424,:476-486,:84-86): out = *flag ? a : b
 * (one int32 decision per step for the whole batch) and its gradient routing (the branch not taken gets zeros) */
int pd_select_rows(const float* a, long lda, const float* b, long ldb, const int* flag, float* out, long ldo, long rows,
                   int cols, void* stream);
int pd_select_rows_bwd(const float* dout, long ldd, const int* flag, float* da, long ldda, float* db, long lddb, long rows,
                       int cols, void* stream);
/* z = mu + sd*eps (eps NULL: z = mu), z row stride ldz;  bwd: dmu = dz, dsd = dz*eps */
int pd_reparam_fwd(const float* mu, const float* sd, const float* eps, int B, int D, float* z, long ldz,
                   void* stream);
int pd_reparam_bwd(const float* dz, long lddz, const float* eps, int B, int D, float* dmu, float* dsd,
                   void* stream);
/* mean over n elements of KL(N(mu,sd) || N(0,1)) */
int pd_kl_fwd(const float* mu, const float* sd, long n, float* out, void* stream);
int pd_kl_bwd(const float* mu, const float* sd, long n, const float* gout, float* dmu, float* dsd, void* stream);

/* ---- optimizer tail (amc_dl/torch_plus/module.py:142-143, scheduler.py:69-74, example.py:4-12): global-norm
 * clip + Adam + exponentially decayed LR with a floor on flat fp32 buffers; norm and step count live on the
 * device.  pd_sumsq_f32 accumulates into out[0]. */
int pd_sumsq_f32(const float* g, long n, float* out, void* stream);
int pd_counter_inc(int* counter, void* stream);
int pd_adam_clip_step(float* p, const float* g, float* m, float* v, long n, const float* sumsq, const int* step,
                      float lr0, float gamma, float lr_min, float b1, float b2, float eps, float clip, void* stream);

/* ---- packed note level of the teacher-forced PianoTree decoder in loss mode (csrc/packed.cu, ops.py "packed notes").
 * Replaces, for training, the dense 15-slot note loop of ptvae.py:370-428 and its heads / duration decoder (:336-368):
 * the reference computes every slot of every (segment, time step) row and lets the loss ignore the PAD targets
 * (ptvae.py:498-511); here rows are sorted by token count (as pack_padded_sequence does for the summary GRU,
 * ptvae.py:446-453), note-level buffers are slot-major (slot n, sorted row r), and a DEVICE table says how many rows of
 * each slot are live.  Kernels are launched for the full extent and skip dead tiles, so one CUDA graph serves every batch.
 * table (64 int32): [0,17) c[t] = rows with more than t tokens; [17,34) cp[t] = min(R, c[t] rounded up to 128);
 * [34,51) 6*cp[t].  A live-row predicate is (cp pointer, slot_rows): row q of a slot-major buffer is live iff
 * q % slot_rows < cp[q / slot_rows]. */
int pd_pack_order(const int* lengths, int R, int* perm, int* inv, int* table, void* stream);
int pd_pack_grid(const int* tok, const int* lengths, const int* perm, int R, int* tok_s, int* pitch_tgt_s, int* dur_tgt_s,
                 int* lengths_s, void* stream);
/* dst (R,C) = src[idx] */
int pd_gather_rows_f32(const float* src, long lds, const int* idx, long R, int C, float* dst, long ldd, void* stream);
/* out (R,C) = sum over slots t < T with r < cp[t] of X (T,R,C) */
int pd_sum_slots_rows_f32(const float* X, long ldt, long ldr, int T, const int* cp, float* out, long ldo, long R, int C,
                          void* stream);
/* column sums over the live rows */
int pd_colsum_rows_f32(const float* X, long ldx, long M, int N, float* out, int accumulate, const int* cp, int slot_rows,
                       void* stream);
/* pd_gemm_tf32 with a live-row predicate: pred 1 = on the rows of A / C (dead 128-row tiles are neither loaded, multiplied
 * nor stored), pred 2 = on the contraction index (weight gradients: only live 32-row k-blocks are accumulated) */
int pd_gemm_tf32_rows(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                      const float* bias, int M, int N, int K, int accumulate, int pred, const int* cp, int slot_rows,
                      void* stream);
/* pd_gru_step_tmax / pd_gru_gates_bwd over the first *nrows rows (a device count) */
int pd_gru_step_tmax_rows(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* x, long ldx,
                          const float* w_x, long ldwx, int K2, const float* b_hh, const float* gi2, long ldgi2, float* hout,
                          long ldho, float* rzn, long ldrzn, float* hn, long ldhn, int B, int H, const int* nrows,
                          void* stream);
int pd_gru_gates_bwd_rows(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3, long lddh3,
                          const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp,
                          float* dgi, long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, int B, int H,
                          const int* nrows, void* stream);
/* duration decoder / embedding gradient with the live-row predicate (TF32 training mode) */
int pd_dur_decode_fwd_rows(const float* h0, long ldh0, long Q, const float* w_ih, const float* b_ih, const float* w_hh,
                           const float* b_hh, const float* sos, const float* w_out, const float* b_out, float* logits,
                           float* S, const int* cp, int slot_rows, void* stream);
int pd_dur_decode_bwd_rows(const float* S, const float* dlogits, long Q, const float* w_ih, const float* b_ih,
                           const float* w_hh, const float* b_hh, const float* sos, const float* w_out, const float* b_out,
                           float* GX, float* dh0, long lddh0, const int* cp, int slot_rows, void* stream);
int pd_note_embed_bwd_rows(const int* tok, long R, const float* g, long ldg, float* dWT, float* dbias, const int* cp,
                           int slot_rows, void* stream);

/* pd_gru128_bwd for the note summarisers.  cp (nullable): length-sorted rows whose dgi is a slot-major slab with skipping
 * consumers -- masked (row, step) entries of dgi are zero-filled only for rows < cp[t].  dout_step >= 0: dout is the (R,128)
 * gradient of that step's output alone (-1: dout is (R,T,128)). */
int pd_gru128_bwd_rows(const float* dout, long dr, long dt, const float* h_all, long hr, long ht, const float* rzn, long zr,
                       long zt, const float* hn, long nr, long nt, const int* lengths, const float* w_hh, float* dgi, long gr,
                       long gt, float* dgh, long qr, long qt, long R, int T, int reverse, const int* cp, int dout_step,
                       void* stream);

/* texture front end, training form: the forward also writes the pooled arg-max position of every output (int8, -1 = ReLU
 * inactive; same layout as out), the backward reads it instead of recomputing the convolution */
int pd_texture_frontend_fwd_ix(const float* pr_mat, const float* w, const float* bias, int B, int C, float* out,
                               signed char* amax, void* stream);
int pd_texture_frontend_bwd_ix(const float* pr_mat, const signed char* amax, int B, int C, const float* gout, float* dw,
                               float* dbias, void* stream);

/* bf16-operand step of the batch-sized recurrences (BASELINE configs[1] "bf16 / fp32-accumulate"): hb_prev / wb are bf16
 * copies of h_prev (B x H) and W_hh (3H x H) (strides in elements); accumulators, gate math, the h_prev of the blend and all
 * saved arrays are fp32; hb_out receives the bf16 copy of the new state.  pd_gru_gates_bwd_zb: pd_gru_gates_bwd_z that also
 * writes a bf16 copy of dgh, the A operand of the bf16 dgh.W_hh GEMM (pd_gemm_bf16). */
int pd_gru_step_tma_bf16(const void* hb_prev, long ldhbp, const void* wb, long ldwb, const float* b_hh, const float* gi,
                         long ldgi, const float* gi2, long ldgi2, const float* hprev, long ldhp, float* hout, long ldho,
                         void* hb_out, long ldhbo, float* rzn, long ldrzn, float* hn, long ldhn, int B, int H, void* stream);
/* the same step with a chosen tile width: units = 32 (one wave of 128 CTAs for a 512-row recurrence) or 64 (64 CTAs: two
 * independent recurrences run side by side on disjoint SMs) */
int pd_gru_step_tma_bf16_units(const void* hb_prev, long ldhbp, const void* wb, long ldwb, const float* b_hh, const float* gi,
                               long ldgi, const float* gi2, long ldgi2, const float* hprev, long ldhp, float* hout, long ldho,
                               void* hb_out, long ldhbo, float* rzn, long ldrzn, float* hn, long ldhn, int B, int H, int units,
                               void* stream);
int pd_gru_gates_bwd_zb(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3, long lddh3,
                        const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp, float* dgi,
                        long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, const int* lengths, int t, int B, int H,
                        float* zero_out, long ldzo, void* dgh_b, long lddghb, void* stream);

/* pd_gru_gates_bwd that also clears zero_out (B,H): the accumulator of the split-K dgh.W_hh GEMM that follows, which then
 * runs with accumulate = 1 and no zero-fill node of its own.  pd_gemm_tf32_splits: the number of K splits pd_gemm_tf32
 * uses for a problem (returns the count, not a status). */
int pd_gru_gates_bwd_z(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3, long lddh3,
                       const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp, float* dgi,
                       long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, const int* lengths, int t, int B, int H,
                       float* zero_out, long ldzo, void* stream);
int pd_gemm_tf32_splits(int M, int N, int K);

/* Library-wide switch: launch the kernels of the recurrent chain (fused step, gate gradients, the tcgen05 GEMM) with the
 * programmatic-dependent-launch attribute, so that a kernel's prologue overlaps its predecessor's tail (csrc/common.cuh).
 * 0 (default) = ordinary launches. */
int pd_set_pdl(int on);

/* ---- gradient exchange over NVLink peer memory (csrc/allreduce_p2p.cu).  Replaces the gradient averaging of the
 * reference's nn.DataParallel replicas (amc_dl/torch_plus/module.py:67-68, 152-157) for one-process-per-GPU training:
 * ONE kernel per gradient bucket gathers the bucket from the per-parameter gradient buffers, exchanges it with the
 * peers (two-shot all-reduce: each rank reduces its 1/W slice with peer loads and stores the result into every copy),
 * scales by `scale` (1/W) and optionally leaves the squared-norm partials of the clip.
 * A rank's "region" = pd_ar_flag_bytes(n_buckets) bytes of flags followed by the flat fp32 buckets, allocated with
 * pd_ipc_alloc (cudaMalloc + zero fill + 64-byte cudaIpcMemHandle_t) and mapped by every peer with pd_ipc_open.
 * pd_allreduce_p2p: peers[p] = rank p's region as mapped here (own region for p == rank); off / n: element offset and
 * count of the bucket in the data part (multiples of 4); bucket: the flag slot / epoch row of this exchange (exchanges of
 * different buckets may overlap on different streams; those of one bucket are issued in the same order on every rank);
 * epoch: n_buckets x pd_ar_limit(1) zeroed uint32 of this rank; err: set to 1 if a peer did not answer within ~20 s (the
 * kernel then gives up instead of hanging); with_norm: leave the squared-norm partials of the averaged bucket
 * (pd_ar_norm_total sums the first n_slots buckets' partials into out[0], same bits on every rank); nblocks <=
 * pd_ar_limit(1), identical on all ranks; src / src_off / src_n / n_src: the gather table (n_src = 0: the bucket is
 * already in place). */
int pd_ar_flag_bytes(int n_buckets);
int pd_ar_limit(int which);
int pd_ipc_alloc(long bytes, void** ptr, void* handle64);
int pd_ipc_open(const void* handle64, void** ptr);
int pd_ipc_close(void* ptr);
int pd_ipc_free(void* ptr);
int pd_allreduce_p2p(const void* const* peers, int rank, int world, long flag_bytes, long off, long n, float scale,
                     void* epoch, int* err, int bucket, int with_norm, int nblocks, const void* const* src,
                     const long* src_off, const long* src_n, int n_src, void* stream);
int pd_ar_norm_total(const void* region, int n_slots, int world, int nblocks, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POLYDIS_B200_H */
