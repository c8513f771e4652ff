/* vcof.h — C ABI of libvcof: the B200-native (sm_100a) kernels behind the VideoCoF
 * denoising hot path (Wan-2.1 DiT block stack + 3D causal VAE).
 *
 * The reference (knightyxp/VideoCoF) has no FFI / operator registry: its boundary is the
 * Python class API of videox_fun.models / videox_fun.pipeline.  This header
 * is the layer UNDER that API: one extern "C" entry per kernel family, plain pointers and
 * sizes only, no torch types.  Each entry cites the reference call site it replaces
 * (paths relative to the reference repo).  videocof_b200/_lib.py binds it with ctypes; the
 * reference-side stub a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; tensors are row-major;
 *     ld* are leading dimensions in ELEMENTS
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered, never
 *     synchronise and never allocate
 *   - return 0 on success, negative on error; vcof_last_error() returns the message
 *     (thread-local).  Nothing here falls back to a CPU path.
 */
#ifndef VCOF_H_
#define VCOF_H_

#ifdef __cplusplus
extern "C" {
#endif

#define VCOF_ABI_VERSION 1

const char* vcof_last_error(void);
int vcof_abi_version(void);

/* ---- GEMM epilogues -------------------------------------------------------------------- */
#define VCOF_EPI_BIAS_BF16 0         /* out_bf16 = bf16(acc + bias)                          */
#define VCOF_EPI_BIAS_GELU_BF16 1    /* out_bf16 = bf16(gelu_tanh(bf16(acc + bias)))         */
#define VCOF_EPI_BIAS_GATE_RES_F32 2 /* out_f32 += gate[n] * bf16(acc + bias)  (gate NULL=1) */
#define VCOF_EPI_BIAS_F32 3          /* out_f32  = float(bf16(acc + bias))                   */
#define VCOF_EPI_RAW_F32 4           /* out_f32  = acc + bias (unrounded; attention scores)   */
#define VCOF_EPI_GATE_ACCUM_BF16 5   /* out_bf16 = bf16(float(out_bf16) + gate[n] * acc): in-place LoRA merge
                                        W += multiplier * alpha/rank * up @ down (utils/lora_utils.py:482-496) */

#define VCOF_EPI_MUL_BF16 6          /* out_bf16 = bf16(float(out_bf16) * bf16(acc + bias)): gated FFN of the text
                                        encoder, `out` holds gelu(gate(x)) (wan_text_encoder.py:129)             */
#define VCOF_EPI_ADD_BF16 7          /* out_bf16 = bf16(float(out_bf16) + bf16(acc + bias)): bf16 residual stream
                                        of the text encoder (wan_text_encoder.py:156-157)                        */

/* Optional flag OR-ed into `epilogue`: tile the output 128 columns wide instead of 256.  For skinny problems (the text
 * encoder's M = 512 tokens) whose 256-wide tiling would occupy less than the machine or spill a few tiles into a second
 * wave; the caller decides (videocof_b200/ops.py: gemm(..., narrow=True)).  Results are identical. */
#define VCOF_GEMM_TILE128 0x100

/* D[M,N] = A[M,K] (bf16) x W[N,K]^T (bf16, nn.Linear layout) with fused epilogue; tcgen05 +
 * TMEM + TMA.  Replaces nn.Linear q/k/v/o, ffn.0/ffn.2, text_embedding, head.head and the
 * patch-embedding Conv3d-as-GEMM: wan_transformer3d.py:264-267, 284-290, 303-304, 457-459,
 * 543, 662-666, 870; the fused epilogues replace :458 (GELU), :499/:504/:511 (gate+residual). */
int vcof_gemm_bf16(const void* a, long long lda, const void* w, long long ldw, const void* bias,
                   const float* gate, void* out, long long ldo, int M, int N, int K, int epilogue,
                   void* stream);

/* out[Lq, heads*128] = softmax(Q K^T * scale) V per head, non-causal, keys [0, kv_len).
 * q/k/(v) are [L, heads*128] bf16 (heads interleaved along the row, as .view(b,s,n,d));
 * with v_transposed != 0, v is V^T stored [heads*128, ldv] (kv contiguous).
 * Replaces attention()/flash_attention(): attention_utils.py:43-149, 152-210 as called from
 * wan_transformer3d.py:294-299 (self) and :325-330 (cross, kv_len = 512). */
int vcof_attn_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                  long long ldv, void* out, long long ldo, int Lq, int Lk, int kv_len, int heads,
                  int head_dim, float softmax_scale, int v_transposed, void* stream);

/* out_bf16[L,C] = ((LayerNorm(x_f32[L,C], eps) * ln_w + ln_b) * (1 + scale) + shift); every
 * per-channel vector is fp32 [C] and may be NULL (identity).  Replaces WanLayerNorm + AdaLN
 * modulate + cast: wan_transformer3d.py:233-243, 495-496, 504 (norm3), 507-508, 547 (head). */
int vcof_ln_modulate(const float* x, long long ldx, const float* ln_w, const float* ln_b,
                     const float* shift, const float* scale, void* out, long long ldo, int L, int C,
                     float eps, void* stream);

/* In place on x_bf16[L,C]: WanRMSNorm over the full row (all heads), then 3-axis RoPE on
 * interleaved pairs.  y = bf16(bf16(x * bf16(rsqrt(mean(x^2)+eps))) * w); if rope_table != NULL
 * token (row_offset + l) of the (F,H,W) grid is rotated by table[pos][i] = (cos, sin), i in
 * [0,64): pairs [0,n_t) use the temporal position tpos[f], [n_t, n_t+n_h) the row, the rest the
 * column.  Rows whose global token id >= F*H*W are normalised but not rotated.
 * Replaces wan_transformer3d.py:214-230 (WanRMSNorm) and :135-211 (rope_apply, incl. the
 * chain-of-frames temporal positions :153-198). */
int vcof_rmsnorm_rope(void* x, long long ldx, const void* weight, float eps, int L, int C,
                      int head_dim, const float* rope_table, const int* tpos, int F, int H, int W,
                      int n_t, int n_h, int row_offset, void* stream);

/* Same arithmetic as vcof_rmsnorm_rope, out of place into the column-blocked layout
 * y[C / cols_per_block][L][cols_per_block] (blocks block_stride elements apart): the send buffer of the
 * sequence-parallel head exchange (one block of heads per destination rank), so the pack costs no extra pass.
 * New (the reference's xfuser all-to-all lives in the absent yunchang package, dist/wan_xfuser.py:98). */
int vcof_rmsnorm_rope_blocked(const void* x, long long ldx, void* y, int cols_per_block, long long block_stride,
                              const void* weight, float eps, int L, int C, int head_dim, const float* rope_table,
                              const int* tpos, int F, int H, int W, int n_t, int n_h, int row_offset, void* stream);

/* Copy between a row-major bf16 [rows, C] matrix (pitch ld) and its column-blocked form
 * [C / cols_per_block][rows][cols_per_block]; to_blocked != 0 packs, 0 unpacks.  Pack / unpack of the head
 * exchange for tensors that have no normalisation pass (V, and the attention output on the way back). */
int vcof_copy_blocked(void* rowmajor, long long ld, void* blocked, long long block_stride, long long rows, int C,
                      int cols_per_block, int to_blocked, void* stream);

/* ---- push-style head exchange (opt-in, VCOF_SP_MODE=push): the producing kernel's stores ARE the transfer --------
 * The three entries below write into up to 16 destination slabs given as device pointers; under sequence parallelism
 * the slabs are the other ranks' receive buffers mapped through NVLink peer memory (torch symmetric memory), so the
 * all-to-all of the head exchange needs no collective call: it overlaps the producing kernel store by store, and a
 * cross-GPU barrier on the stream orders it against the consumer.  New design (the reference's exchange is xfuser /
 * yunchang's NCCL all-to-all, dist/wan_xfuser.py:98; absent package).  The pointer array itself is a HOST array. */

/* vcof_rmsnorm_rope with the result scattered: block b = columns [b*C/n_blocks, (b+1)*C/n_blocks) of every row goes to
 * the dense [L, C/n_blocks] slab block_ptrs[b]; x is not modified. */
int vcof_rmsnorm_rope_scatter(const void* x, long long ldx, void* const* block_ptrs, int n_blocks, const void* weight,
                              float eps, int L, int C, int head_dim, const float* rope_table, const int* tpos, int F,
                              int H, int W, int n_t, int n_h, int row_offset, void* stream);

/* Column blocks of a row-major bf16 [rows, C] matrix (pitch ld) to the slabs block_ptrs[b] ([rows, C/n_blocks]). */
int vcof_copy_scatter(const void* rowmajor, long long ld, void* const* block_ptrs, int n_blocks, long long rows, int C,
                      void* stream);

/* Row chunks of a bf16 [n_chunks*rows, cols] matrix (pitch ld) to the slabs chunk_ptrs[c] ([rows, cols]): the unfused
 * form of the return leg (attention into a local buffer, then this copy), kept for A/B against vcof_attn_fwd_scatter. */
int vcof_copy_rows_scatter(const void* src, long long ld, void* const* chunk_ptrs, int n_chunks, long long rows,
                           int cols, void* stream);

/* vcof_attn_fwd (natural V layout) with the output rows scattered: chunk c = query rows [c*rows_per_chunk,
 * (c+1)*rows_per_chunk) is stored to the dense [rows_per_chunk, ldo] slab out_chunks[c] (HOST array of n_chunks <= 16
 * device pointers).  Return leg of the push exchange: chunk c is the rows rank c owns and out_chunks[c] its receive
 * buffer, so the attention epilogue's own stores carry the output over NVLink — compute and transfer in one kernel. */
int vcof_attn_fwd_scatter(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                          void* const* out_chunks, int n_chunks, int rows_per_chunk, long long ldo, int Lq, int Lk,
                          int kv_len, int heads, int head_dim, float softmax_scale, void* stream);

/* Patchify latents x_bf16[Cin, F, H, W] -> tokens a_bf16[F*(H/2)*(W/2), Cin*4], column order
 * (c, ph, pw) = the flattened Conv3d weight [C, Cin, 1, 2, 2].  wan_transformer3d.py:870, 879. */
int vcof_patchify(const void* x, void* a, int Cin, int F, int H, int W, void* stream);

/* Unpatchify head output y_bf16[L, 4*Cout] (column order (ph, pw, c)) -> out_bf16[Cout, F, H, W]
 * (H, W are the LATENT sizes, L = F*(H/2)*(W/2)).  wan_transformer3d.py:1108-1131. */
int vcof_unpatchify(const void* y, long long ldy, void* out, int Cout, int F, int H, int W,
                    void* stream);

/* Small fp32 linear for the timestep path (runs under autocast(fp32) in the reference):
 * out_f32[B,N] = act_out( act_in(x_f32[B,K]) @ W_bf16[N,K]^T + bias_bf16[N] ), act: 0 none, 1 SiLU.
 * wan_transformer3d.py:668-670, 913-929. */
int vcof_linear_f32(const float* x, const void* w, const void* bias, float* out, int B, int N, int K,
                    int act_in, int act_out, void* stream);

/* ---- 3D causal VAE (channels-last bf16 activations [T, H, W, C]) ------------------------- */

/* Implicit-GEMM convolution on tcgen05, im2col-free: per filter tap one TMA box of the shifted input
 * patch feeds the MMA directly; TMA zero-fill provides the spatial and causal-temporal padding.
 *   x_dims[5]    (c_inner, W, P, H, T) of the input view (P = 1, or 2 for the stride-2 "parity" view
 *                of the down-sampler where c_inner = 2*Cin); x_strides[4]: element strides of dims 1..4
 *   taps[5*i..]  (c_base, dw, p, dh, dt): coordinate offsets of tap i added to the tile origin; taps come in groups
 *                of `tgroup` (1 or 3) consecutive entries that differ only by dt, dt+1, dt+2 — one TMA box with a
 *                t-extent of tgroup feeds the whole group (the TMA unit's cost is per box)
 *   w            bf16 [k_total/kc, n_total, kc]: slice ((group * cin/kc + chunk) * tgroup + j) holds the kc input
 *                channels `chunk` of tap (group, j) for every output channel (n_total padded to a multiple of 16);
 *                cin is the per-tap K extent, zero-padded in the weights to a multiple of kc (the activation box may
 *                read past the tensor's channels: TMA zero-fills, or the zero weights cancel a neighbour's data)
 *   geom[17]     T_out, H_out, W_out, t_stride, n_total, n_tile, ot_mul, ot_add, oh_mul, oh_add, ow_mul,
 *                ow_add, Hs, Ws, interleave_half, n_store, kc (32 | 64 channels per K slice = 64 | 128-byte TMA
 *                rows; 64 whenever the layer has >= 64 input channels) — output position (t,h,w) is stored at
 *                [t*ot_mul+ot_add, h*oh_mul+oh_add, w*ow_mul+ow_add] of a [*, Hs, Ws, ldc] tensor;
 *                interleave_half > 0 sends channels >= half to the next frame (upsample3d, wan_vae.py:137-141)
 *   bias fp32 [n_total] | NULL;  residual bf16 (same addressing as out) | NULL;  clamp > 0 clamps.
 *   act_out | NULL: additionally stores silu(rms_norm(result) * act_gamma) — the RMS_norm + SiLU that
 *                opens the NEXT layer (wan_vae.py:197-201) — with the same addressing; `out` may then be NULL.
 * Replaces CausalConv3d / Conv2d of wan_vae.py:21-40, 80-100, 107-163, 190-224, 318-320, 423-425. */
int vcof_conv_igemm(const void* x, const long long* x_dims, const long long* x_strides, const void* w,
                    int k_total, const short* taps, int ntaps, int tgroup, int cin, const int* geom, const float* bias,
                    const void* residual, void* out, long long ldc, float clamp, void* act_out,
                    const float* act_gamma, void* stream);

/* Line-resident kernel (videocof_b200/vae.py routes layers of up to 128 output channels here; VCOF_CONV_LINES=0|1
 * overrides): the same stride-1 'same' 3x3 / 3x3x3 convolution as vcof_conv_igemm on channels-last bf16 [T, H, W, C], with input lines
 * and per-phase weight tiles kept resident in shared memory (csrc/conv_sm100.cu, conv_lines_kernel).
 *   x, x_dims, x_strides   the plain 5-D view (C, W, 1, H, T) as for vcof_conv_igemm
 *   w            bf16 [cin/32 * kt * 9, n_total, 32]: slice ((chunk * kt + dt) * 3 + dh) * 3 + dw holds the 32 input
 *                channels `chunk` of tap (dt, dh, dw) for every output channel
 *   kt, t0       temporal taps (1 | 3) and the input frame of the first one relative to the output frame
 *                (-(kt-1) for the causal convolution, plus the halo shift under temporal sharding)
 *   geom[7]      T_out, H_out, W_out, n_total, n_tile (channels per pass, <= 256), rows (output rows per work item,
 *                1..4 and <= min(5, 512 / roundup32(n_tile)), the accumulators of the kernel's TMEM ring), n_store
 * bias / residual / out / ldc / clamp / act_out / act_gamma as vcof_conv_igemm (plain output addressing).
 * Replaces CausalConv3d / Conv2d of wan_vae.py:21-40, 190-224 for the 3x3(x3) stride-1 layers. */
int vcof_conv_lines(const void* x, const long long* x_dims, const long long* x_strides, const void* w, int cin, int kt,
                    int t0, const int* geom, const float* bias, const void* residual, void* out, long long ldc,
                    float clamp, void* act_out, const float* act_gamma, void* stream);

/* y = [silu]( x / max(||x||_2, 1e-12) * sqrt(C) * gamma ) per position, channels-last; RMS_norm (+ nn.SiLU)
 * of wan_vae.py:43-58, fp32 intermediates and one bf16 rounding at the store (the reference's ATen chain rounds
 * after every op). */
int vcof_rms_silu_cl(const void* x, long long ldx, const float* gamma, void* y, long long ldy,
                     long long npos, int C, int silu, void* stream);

/* [C, T*H*W] bf16 -> channels-last [T*H*W, Cp] (extra channels zero), optional x/div[c]+add[c]
 * (latent de-normalisation z / (1/std) + mean, wan_vae.py:553-558). */
int vcof_nchw_to_cl(const void* x, void* y, int C, int Cp, long long thw, const float* div, const float* add,
                    void* stream);

/* channels-last [T*H*W, ldx] -> [C, T*H*W] bf16, optional (x - sub[c]) * mul[c] (mu normalisation,
 * wan_vae.py:540-546). */
int vcof_cl_to_nchw(const void* x, long long ldx, void* y, int C, long long thw, const float* sub,
                    const float* mul, void* stream);

/* Fused single-head attention of the VAE's AttentionBlock (wan_vae.py:244-266: per frame,
 * F.scaled_dot_product_attention over h*w tokens with d = C = 384).  qkv bf16 [T, N, >= 3C] with row pitch ld: the
 * to_qkv output, q = columns [0, C), k = [C, 2C), v = [2C, 3C) (:251-256); out bf16 [T, N, C] with row pitch ldo:
 * out[t] = softmax(q[t] k[t]^T * softmax_scale) v[t].  Scores and probabilities stay in tensor / shared memory.
 * C must be 384. */
int vcof_vae_attn(const void* qkv, long long ld, void* out, long long ldo, int T, int N, int C, float softmax_scale,
                  void* stream);

/* p_bf16[rows, n] = softmax(s_f32[rows, n] * scale) — the unfused form of the VAE attention (GEMM with raw fp32
 * scores -> this -> GEMM), kept as the cross-check of vcof_vae_attn (VCOF_VAE_ATTN=unfused). */
int vcof_softmax_rows(const float* s, long long lds, void* p, long long ldp, int rows, int n, float scale,
                      void* stream);

/* ---- umT5 text encoder (bias-free Linears run on vcof_gemm_bf16) -------------------- */

/* out_bf16[i, :] = table_bf16[ids[i], :] for i < n; ids are int64 DEVICE values in [0, vocab) (an id outside the range
 * yields a zero row, never a wild read; the host wrapper validates ids).  nn.Embedding lookup,
 * wan_text_encoder.py:269-270, 285. */
int vcof_embed_rows(const long long* ids, const void* table, long long ldt, long long vocab, void* out, long long ldo,
                    long long n, int C, void* stream);

/* y = bf16(w * bf16(x * rsqrt(mean(x^2) + eps))) per row of x_bf16[rows, C]; the fp32 factor is not rounded
 * (T5LayerNorm, wan_text_encoder.py:45-57 — differs from WanRMSNorm, which rounds it). */
int vcof_t5_rmsnorm(const void* x, long long ldx, const void* weight, void* y, long long ldy, long long rows, int C,
                    float eps, void* stream);

/* T5 self-attention for B samples of L <= 512 tokens: q/k/v/out are bf16 [B*L, heads*head_dim] (sample-major rows,
 * heads interleaved along the row as .view(b, -1, n, c)); scores = q.k (NO 1/sqrt(d)) + bias_rel[h][(j - i) + L - 1]
 * with bias_rel fp32 [heads, bias_ld >= 2L-1] (the position bias depends on key - query only); a key with
 * key_mask[b*L + j] == 0 gets finfo(bf16).min in place of its bias (masked_fill_, :94-98; key_mask int32 or NULL);
 * fp32 softmax, probabilities rounded to bf16, fp32 P.V, bf16 store.  head_dim in {16, 32, 64, 128}.
 * Replaces T5Attention.forward + T5RelativeEmbedding.forward, wan_text_encoder.py:76-112, 207-222. */
int vcof_t5_attn(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                 long long ldo, const float* bias_rel, int bias_ld, const int* key_mask, int B, int L, int heads,
                 int head_dim, void* stream);

/* ---- frame bytes at the two ends of the pipeline ------------------------------------ */

/* out_u8[npos, C] = trunc(255 * clamp(bf16(bf16(x / 2) + 0.5), 0, 1)) for the first C channels of channels-last bf16
 * x[npos, ldx] (the decoder output, C = 3): WanPipeline.decode_latents' bf16 (frames / 2 + 0.5).clamp(0, 1)
 * (pipeline_wan.py:425-426) + .cpu().float() (:427) + save_videos_grid's (x * 255).astype(uint8)
 * (videox_fun/utils/utils.py:59-68) in one pass; [T, H, W, 3] byte frames come out ready for the encoder. Bit-exact. */
int vcof_cl_to_u8(const void* x, long long ldx, unsigned char* out, long long npos, int C, void* stream);

/* y_bf16[npos, Cp] (channels-last, channels >= C zero) = bf16(fp32(u) * fp32(2/255) - 1) from byte frames
 * frames[npos, C]: load_video_frames' scaling (fast_infer.py:86-88) + the cast to the VAE dtype (pipeline_wan.py:397)
 * + the NCHW -> channels-last pass of the encoder's first layer.  Bit-exact. */
int vcof_u8_to_cl(const unsigned char* frames, void* y, long long npos, int C, int Cp, void* stream);

/* ---- diagnostics ---------------------------------------------------------------------- */
/* The per-element functions of vcof_cl_to_u8 / vcof_u8_to_cl evaluated on the HOST over host arrays (bf16 values as
 * raw 16-bit patterns): the CPU test suite pins the kernels' arithmetic exhaustively.  No product code calls these. */
int vcof_debug_frame_u8_host(const unsigned short* host_bf16_bits, unsigned char* host_out, long long n);
int vcof_debug_video_bf16_host(const unsigned char* host_bytes, unsigned short* host_bf16_bits, long long n);

#ifdef __cplusplus
}
#endif
#endif /* VCOF_H_ */
