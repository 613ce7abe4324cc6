/*
 * ivit_b200.h -- C ABI of the B200 (sm_100a) integer-only ViT operator library.
 *
 * This is the drop-in boundary for the hot path of zkkli/I-ViT: the forward of the
 * seven operator classes exported by the reference's
 *     models/quantization_utils/__init__.py:1
 * (QuantLinear, QuantAct, QuantConv2d, QuantMatMul, IntLayerNorm, IntSoftmax, IntGELU)
 * plus the primitives of models/quantization_utils/quant_utils.py they call.
 * Each entry point below names the reference function it replaces (file:line, relative
 * to the reference checkout).  The reference is pure Python/PyTorch, so the "FFI" a
 * maintainer would add is a ctypes stub (shown in INTEGRATION.md; the one this repo
 * ships is i-vit_b200/_lib.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked HOST
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), takes the
 *     context created by ivit_create(), returns 0 on success or a negative IVIT_E* code;
 *     ivit_last_error() returns a thread-local message.  No exceptions cross the boundary.
 *   - integer tensors are row-major [rows, cols] in their narrowest storage type
 *     (IVIT_I8 / IVIT_I16 / IVIT_I32); "carrier" tensors are the reference's fp32
 *     integer*scale representation (IVIT_F32)
 *   - dyadic multipliers are ivit_dyadic_t {m, e}:  out = RNE(z * m / 2^e)
 *     with m int32 (|m| in [2^30, 2^31)) and e in [-1, 63]; length 1 (scalar) or cols
 *   - a thread may use one context at a time (thread-compatible, not thread-safe)
 */
#ifndef IVIT_B200_H
#define IVIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define IVIT_API
#else
#define IVIT_API __attribute__((visibility("default")))
#endif

typedef struct ivit_ctx ivit_ctx;
typedef void* ivit_stream;

typedef struct ivit_dyadic_t {
    int32_t m;
    int32_t e;
} ivit_dyadic_t;

enum { IVIT_I8 = 0, IVIT_I16 = 1, IVIT_I32 = 2, IVIT_F32 = 3, IVIT_U8 = 4, IVIT_F64 = 5 };

enum {
    IVIT_OK = 0,
    IVIT_EINVAL = -1,   /* bad argument (shape / dtype / alignment / null pointer)           */
    IVIT_ECUDA = -2,    /* CUDA runtime / driver error (message has the CUDA error string)    */
    IVIT_ENOTSUP = -3,  /* valid in the reference but outside this library's supported domain */
    IVIT_ENODEV = -4    /* no sm_100 device                                                   */
};

/* ---- library / context ----------------------------------------------------------- */
IVIT_API int ivit_version(void);
IVIT_API const char* ivit_last_error(void);
/* Creates a context on CUDA device `device` (must be compute capability 10.x). */
IVIT_API int ivit_create(int device, ivit_ctx** out);
IVIT_API int ivit_destroy(ivit_ctx* ctx);
IVIT_API int ivit_num_sms(ivit_ctx* ctx);

/* ---- quant_utils.py primitives ---------------------------------------------------- */

/* batch_frexp (quant_utils.py:150-175) of fp64(s_in[i]) / fp64(fp32(*s_out))
 * (the ratio formed in fixedpoint_mul.forward, quant_utils.py:221-228), on the device,
 * no host round trip: m = round_half_away(mant * 2^31), e = 31 - exp, normalised so
 * that m fits int32, e clamped to [-1, 63].  s_in: n floats; s_out: 1 float (device). */
IVIT_API int ivit_dyadic(ivit_ctx*, const float* s_in, int n, const float* s_out,
                         ivit_dyadic_t* out, ivit_stream stream);

/* SymmetricQuantFunction.forward + linear_quantize (quant_utils.py:12-48, 77-96):
 * q = clamp(RNE(fp32(1/s) * x), -2^(b-1), 2^(b-1)-1); out_dtype IVIT_I8 / IVIT_I16 / IVIT_I32.
 * scale: device, ns entries; element i uses scale[(i / inner) % ns] (ns = 1: scalar;
 * per-row weight scales: ns = rows, inner = row length). */
IVIT_API int ivit_quantize_f32(ivit_ctx*, const float* x, int64_t n, const float* scale,
                               int64_t ns, int64_t inner, int bits, int out_dtype, void* out,
                               ivit_stream stream);

/* Carrier -> integer: z = RNE(x / s[c]) (first line of fixedpoint_mul.forward,
 * quant_utils.py:220; also the x / scaling_factor of quant_modules.py:94,224-225,359,426,484).
 * x: [rows, cols] carrier, x_dtype IVIT_F32 or IVIT_F64 (an fp64 carrier is what keeps
 * IntLayerNorm outputs, up to 2^30, exact -- the reference's modules return fp64 there when
 * their own input carrier is exact); s: 1 or cols entries; saturates to out_dtype. */
IVIT_API int ivit_carrier_to_int(ivit_ctx*, const void* x, int x_dtype, int64_t rows, int cols,
                                 const float* s, int s_len, int out_dtype, void* out,
                                 ivit_stream stream);

/* Integer -> carrier: x = float(q) * s[c] (the `* scaling_factor` on every operator's
 * return, e.g. quant_modules.py:96-97,206,228,384,445,497). */
IVIT_API int ivit_int_to_carrier(ivit_ctx*, const void* q, int q_dtype, int64_t rows, int cols,
                                 const float* s, int s_len, int out_dtype, void* out,
                                 ivit_stream stream);

/* fixedpoint_mul.forward (quant_utils.py:192-253) on integers:
 *   out = clamp( RNE(z*m/2^e) [+ RNE(w*m1/2^e1)], -2^(bits-1), 2^(bits-1)-1 )
 * z: [rows, cols] (z_dtype); me: me_len (1|cols) entries; optional residual w
 * (w_dtype, [w_rows, cols], w_rows divides rows: periodic broadcast, e.g. pos_embed over the
 * batch, vit_quant.py:264-265) with me1 (me1_len 1|cols).
 * bits in {4, 8, 16, 32}; out_dtype must hold `bits`. */
IVIT_API int ivit_requant(ivit_ctx*, const void* z, int z_dtype, int64_t rows, int cols,
                          const ivit_dyadic_t* me, int me_len,
                          const void* w, int w_dtype, int64_t w_rows,
                          const ivit_dyadic_t* me1, int me1_len,
                          int bits, int out_dtype, void* out, ivit_stream stream);

/* ---- integer contractions --------------------------------------------------------- */

/* Epilogue selector for ivit_gemm_i8. */
enum {
    IVIT_EPI_RAW_I32 = 0,   /* out int32 = acc + bias                                         */
    IVIT_EPI_REQUANT = 1,   /* out = clamp(RNE((acc+bias)*m[n]/2^e[n]) [+ residual])  i8/i16 */
    IVIT_EPI_CARRIER = 2    /* out fp32 = float(acc + bias) * scale[n]                        */
};

typedef struct ivit_gemm_epilogue {
    int mode;                      /* IVIT_EPI_*                                              */
    const int32_t* bias;           /* [N] or NULL                                             */
    const ivit_dyadic_t* me;       /* [N] (per output channel) for IVIT_EPI_REQUANT           */
    int bits;                      /* 8 or 16 for IVIT_EPI_REQUANT                            */
    const void* residual;          /* optional [M, N] residual (res_dtype), IVIT_EPI_REQUANT  */
    int res_dtype;
    int64_t res_ld;                /* leading dimension (elements) of residual                */
    ivit_dyadic_t res_me;          /* scalar dyadic of the residual                           */
    int two_stage;                 /* != 0: q1 = clamp(RNE(z*me[n]), bits); out = clamp(RNE(q1*me2) + RNE(res*res_me), bits)
                                      (a per-channel QuantAct followed by a residual QuantAct, vit_quant.py:85,135) */
    ivit_dyadic_t me2;             /* scalar dyadic of the second stage                       */
    const float* scale;            /* [N] for IVIT_EPI_CARRIER                                */
    int out_dtype;                 /* IVIT_I8 / IVIT_I16 / IVIT_I32 / IVIT_F32                */
    int64_t out_ld;                /* leading dimension (elements) of out, >= N               */
    int acc_bits;                  /* optional bound |acc + bias| < 2^acc_bits (0 = unknown = 31): lets
                                      the epilogue prove that an exact .5 tie of the per-channel requant
                                      is unreachable for most channels (those keep a 3-instruction form) */
} ivit_gemm_epilogue;

/* QuantLinear.forward contraction (quant_modules.py:93-97), also QuantConv2d on unfolded
 * patches (quant_modules.py:325-330):   acc[i,j] = sum_k A[i,k] * W[j,k]
 * A: int8 [M, K] (lda elements between rows), W: int8 [N, K] (row-major, the
 * `weight_integer` buffer), tcgen05 kind::i8 tensor-core path.  K % 16 == 0, lda % 16 == 0. */
IVIT_API int ivit_gemm_i8(ivit_ctx*, const int8_t* A, int64_t lda, const int8_t* W,
                          int64_t M, int64_t N, int64_t K, const ivit_gemm_epilogue* epi,
                          void* out, ivit_stream stream);

/* QuantMatMul.forward contraction (quant_modules.py:223-228), batched, raw int32 result:
 *   C[b] = A[b] (a_dtype, [M,K], row stride lda, batch stride sa) x B[b]
 * trans_b != 0: B[b] is [N,K] row-major (ldb) and C = A B^T; else B[b] is [K,N] (ldb).
 * a_dtype IVIT_I8 or IVIT_I16 (DeiT's 16-bit softmax output, vit_quant.py:54,79);
 * b_dtype IVIT_I8.  General strided form so that q/k/v views of the qkv buffer
 * (vit_quant.py:63-71) need no copies. */
IVIT_API int ivit_bmm_i32(ivit_ctx*, const void* A, int a_dtype, int64_t lda, int64_t sa,
                          const int8_t* B, int64_t ldb, int64_t sb, int trans_b,
                          int64_t batch, int M, int N, int K,
                          int32_t* C, int64_t ldc, int64_t sc, ivit_stream stream);

/* ---- row operators ----------------------------------------------------------------- */

/* IntLayerNorm.forward (quant_modules.py:353-386), integer part:
 *   mu = RNE(sum/C); y = q - mu; V = sum y^2; 10-step integer sqrt from 2^16;
 *   F = floor((2^31-1)/std); out = floor(y*F/2) + bias_int[c]
 * If `me` != NULL the following QuantAct (per-channel dyadic, `bits`) is fused:
 *   out = clamp(RNE(out * m[c] / 2^e[c])).   x: [rows, C] (IVIT_I8/I16/I32). */
IVIT_API int ivit_layernorm(ivit_ctx*, const void* x, int x_dtype, int64_t rows, int C,
                            const int32_t* bias_int, const ivit_dyadic_t* me, int bits,
                            int out_dtype, void* out, ivit_stream stream);

/* IntSoftmax.forward (quant_modules.py:469-497), Shiftmax over the last dim.
 * q: [rows, cols] IVIT_I8 (or IVIT_I32 with values in int8 range); x0 = floor(-1/s) (host,
 * fp32 arithmetic) in [-2^24, -1] (any scale the reference can produce), n = 15, out_bits 16 (out IVIT_I16) or 8
 * (out IVIT_I8); x0 < -65536 (scale below 2^-16: the reference's clamped sum lets the result leave the nominal range,
 * quant_modules.py:491-493) needs out IVIT_I32. */
IVIT_API int ivit_shiftmax(ivit_ctx*, const void* q, int q_dtype, int64_t rows, int cols,
                           int32_t x0, int n, int out_bits, int out_dtype, void* out,
                           ivit_stream stream);

/* IntGELU.forward (quant_modules.py:410-445), ShiftGELU over the last dim (row max).
 * q: [rows, cols] IVIT_I8; x0 = floor(-1/fp32(s*1.702)) in [-2^24, -1]; n = 23; sigmoid bits 8.
 * out = q * sigma (IVIT_I16 while |x0| <= 255, IVIT_I32 always); if `me` != NULL the following scalar QuantAct is
 * fused: out = clamp(RNE(q*sigma*m/2^e), bits) (out IVIT_I8). */
IVIT_API int ivit_shiftgelu(ivit_ctx*, const void* q, int q_dtype, int64_t rows, int cols,
                            int32_t x0, int n, const ivit_dyadic_t* me, int bits,
                            int out_dtype, void* out, ivit_stream stream);

/* Fused integer attention for one QuantLinear(qkv) output (vit_quant.py:63-83;
 * swin_quant.py:128-164):
 *   S = Q K^T -> qact_attn1 (scalar me_s, 8 bit) [-> + rel-pos bias (qact2) -> + mask]
 *     -> Shiftmax (x0, n, p_bits) -> P V -> qact (scalar me_o, 8 bit)
 * qkv: int8 [tokens_total, 3*H*D], token t of sequence b at row b*n_tok + t, column
 * which*H*D + h*D + d.  out: int8 [tokens_total, H*D].  Scores never leave the chip.
 * relbias: optional int8 [H, n_tok, n_tok] already requantised operand of qact2 with
 * me_b (scalar) for the bias and me_s2 for the scores (swin_quant.py:149);
 * mask: optional int32 [n_win, n_tok, n_tok] integer mask addend (RNE(-100/s), App. A.5),
 * window index = b % n_win. */
typedef struct ivit_attn_params {
    int n_seq, n_tok, n_heads, head_dim;
    ivit_dyadic_t me_s;        /* scores requant (qact_attn1)                                */
    int32_t x0;                /* floor(-1/s_attn)                                           */
    int n;                     /* 15                                                         */
    int p_bits;                /* 16 (DeiT) or 8 (Swin)                                      */
    ivit_dyadic_t me_o;        /* P.V requant (attn.qact2 / qact3)                           */
    const int8_t* relbias;     /* optional                                                   */
    ivit_dyadic_t me_s2, me_b; /* qact2(scores, bias)                                        */
    const int32_t* mask;       /* optional                                                   */
    int n_win;
} ivit_attn_params;

IVIT_API int ivit_attention_i8(ivit_ctx*, const int8_t* qkv, const ivit_attn_params* p,
                               int8_t* out, ivit_stream stream);

/* ---- data movement used by the DeiT stem/tail (graph glue, vit_quant.py:254-276) ---- */

/* Unfold non-overlapping p x p patches of an int8 NCHW image into GEMM rows
 * (QuantConv2d with kernel == stride, layers_quant.py:172-177,190-191):
 * out[(b*Hp + i)*Wp + j, (c*p + u)*p + v] = x[b, c, i*p+u, j*p+v]. */
IVIT_API int ivit_patchify_i8(ivit_ctx*, const int8_t* x, int B, int Cin, int H, int W, int p,
                              int8_t* out, ivit_stream stream);

/* ---- hot-path specialisations (same results as the general entry points above) ---------- */

/* IntGELU followed by a scalar 8-bit QuantAct (layers_quant.py:147-148) depends only on
 * (q, rowmax(q)) in int8 x int8.  ivit_shiftgelu_build_lut evaluates the general formula
 * (quant_modules.py:410-445 + fixedpoint_mul) ONCE per layer into a 64 KiB table
 *   lut[(mx+128)*256 + (q+128)] = clamp8(RNE(q * sigma(q, mx) * m / 2^e)),  q <= mx
 * and ivit_shiftgelu_lut applies it: out[r,c] = lut[(rowmax_r+128)*256 + q[r,c]+128].
 * Bit-identical to ivit_shiftgelu(..., me, 8, IVIT_I8, ...).  cols % 16 == 0, cols <= 4096. */
IVIT_API int ivit_shiftgelu_build_lut(ivit_ctx*, int32_t x0, int n, const ivit_dyadic_t* me, int bits,
                                      int8_t* lut, ivit_stream stream);
IVIT_API int ivit_shiftgelu_lut(ivit_ctx*, const int8_t* q, int64_t rows, int cols, const int8_t* lut,
                                int8_t* out, ivit_stream stream);

/* IntLayerNorm + per-channel 8-bit QuantAct on an int16 residual stream (vit_quant.py:131-132,
 * 137-138): vectorised form of ivit_layernorm(x, IVIT_I16, ..., me, 8, IVIT_I8, ...).
 * C % 8 == 0, C <= 1024. */
IVIT_API int ivit_layernorm_i16_i8(ivit_ctx*, const int16_t* x, int64_t rows, int C,
                                   const int32_t* bias_int, const ivit_dyadic_t* me, int8_t* out,
                                   ivit_stream stream);

/* Input QuantAct (vit_quant.py:257, quant_utils.py:48,90-92) fused with the patch unfold of a kernel == stride
 * QuantConv2d (layers_quant.py:190): fp32 NCHW image -> int8 GEMM rows, same result as
 * ivit_quantize_f32(bits 8) followed by ivit_patchify_i8.  patch % 4 == 0. */
IVIT_API int ivit_quantize_patchify(ivit_ctx*, const float* img, const float* scale, int B, int Cin, int H,
                                    int W, int p, int8_t* out, ivit_stream stream);

/* uint8 NCHW image -> int8 GEMM rows in one pass: ToTensor (u / 255), Normalize ((t - mean[c]) / std[c]) -- the
 * reference's eval transform, utils/data_utils.py:90-91 -- then the input QuantAct and the 16 x 16 patch unfold of
 * ivit_quantize_patchify.  Every step is the same correctly-rounded fp32 operation torch performs, so the result equals
 * the reference's pipeline on the decoded uint8 image bit for bit; it is evaluated once per (channel, byte value) into a
 * 768-entry table on the device and applied by lookup.  mean/std: device fp32 [Cin]; patch 16; Cin <= 4.
 * Moves 4x fewer bytes host -> device than the fp32 entry point. */
IVIT_API int ivit_quantize_patchify_u8(ivit_ctx*, const uint8_t* img, const float* mean, const float* std,
                                       const float* scale, int B, int Cin, int H, int W, int p, int8_t* out,
                                       ivit_stream stream);

/* Vectorised form of ivit_embed_tokens (bits 16, C % 8 == 0, dyadic exponents in [16, 62]). */
IVIT_API int ivit_embed_tokens_fast(ivit_ctx*, const int16_t* pe, const int32_t* cls, const int16_t* pos,
                                    int B, int n_tok, int C, ivit_dyadic_t me, ivit_dyadic_t me_res,
                                    int16_t* out, ivit_stream stream);

/* DeiT stem glue: cls-token concatenation followed by the position-embedding residual QuantAct
 * (vit_quant.py:259-265): out[b,t,:] = clamp(RNE(z*me) + RNE(pos[t,:]*me_res), bits) with
 * z = cls (int32 [C], RNE(cls_token / s)) for t == 0 and pe[b, t-1, :] (int16 [B*(n_tok-1), C],
 * the patch_embed.qact output) otherwise.  pos: int16 [n_tok, C] (qact_pos output).  out int16. */
IVIT_API int ivit_embed_tokens(ivit_ctx*, const int16_t* pe, const int32_t* cls, const int16_t* pos,
                               int B, int n_tok, int C, ivit_dyadic_t me, ivit_dyadic_t me_res,
                               int bits, int16_t* out, ivit_stream stream);

/* ---- Swin hot path (swin_quant.py:121-169, 251-301, 328-349, 539-558) -------------------- */

/* Fused window attention on the tcgen05 tensor cores for the shape every Swin of the reference zoo uses (49-token
 * windows, head_dim 32, 8-bit Shiftmax): same arithmetic and results as ivit_attention_i8 with relbias / mask, two
 * windows per 128-row MMA tile.  qkv: int8 [n_win * 49, 3 * 32 * n_heads] (windows contiguous), out: int8
 * [n_win * 49, 32 * n_heads].
 *   bias_rq    int16 [n_heads, 49, 49]: the identity branch of qact2 ALREADY requantised, RNE(bias_int8 * mb / 2^eb)
 *              (static per block: ivit_requant of the gathered relative-position bias, swin_quant.py:142-149)
 *   mask_bits  optional uint64 [n_win_img, 49]: bit j of entry (w, i) set = key j masked for query i in window w of the
 *              image (attn_mask != 0, swin_quant.py:223-247); window index = global window % n_win_img
 *   mask_add   the integer addend of a masked entry, RNE(-100 / s) (App. A.5); only checked: it must saturate Shiftmax
 * Returns IVIT_ENOTSUP (nothing launched) when a scale is outside the fast-form domain; ivit_attention_i8 covers those. */
typedef struct ivit_winattn_params {
    int n_win, n_heads, n_tok, head_dim;
    ivit_dyadic_t me_s;        /* scores requant (qact_attn1)                     */
    ivit_dyadic_t me_s2;       /* qact2, scores branch                            */
    int32_t x0;                /* floor(-1/s_2)                                   */
    int n;                     /* 15                                              */
    int p_bits;                /* 8                                               */
    ivit_dyadic_t me_o;        /* P.V requant (qact3)                             */
    const int16_t* bias_rq;
    const uint64_t* mask_bits;
    int n_win_img;
    int32_t mask_add;
} ivit_winattn_params;

IVIT_API int ivit_window_attention_i8(ivit_ctx*, const int8_t* qkv, const ivit_winattn_params* p, int8_t* out,
                                      ivit_stream stream);

/* IntLayerNorm + per-channel QuantAct (as ivit_layernorm_i16_i8) over GATHERED rows: the window glue of Swin as index
 * math in the row load.  Images hold L_in input rows and L_out output rows each; output row r of an image is built from
 *   G == 1: input row rowmap[r] (int32 [L_out]; NULL = identity)  -- torch.roll + window_partition of this block composed
 *           with window_reverse + roll-back of the previous one (swin_quant.py:259-288).  If xcopy != NULL the gathered
 *           int16 row is also stored there (the residual stream in the new order);
 *   G == 4: the concatenation of input rows rowmap[4r .. 4r+3] (int32 [L_out * 4]), each C/4 wide -- PatchMerging's strided
 *           2x2 gather + cat (swin_quant.py:337-341) feeding norm (:344).
 * x: int16 [images * L_in, C / G]; out: int8 [rows_out, C]; C % (8 G) == 0, C <= 2048. */
IVIT_API int ivit_layernorm_gather_i16_i8(ivit_ctx*, const int16_t* x, int64_t rows_out, int C, int G,
                                          const int32_t* rowmap, int L_out, int L_in, const int32_t* bias_int,
                                          const ivit_dyadic_t* me, int8_t* out, int16_t* xcopy, ivit_stream stream);

/* Swin patch embedding tail (layers_quant.py:193-195, swin_quant.py:546): IntLayerNorm over the 8-bit qact_before_norm
 * output followed by TWO 16-bit QuantActs, patch_embed.qact (per channel, me[C]) and the model's qact1 (scalar me2):
 * out = clamp16(RNE(clamp16(RNE(LN(x) * m[c] / 2^e[c])) * m2 / 2^e2)).  x: int8 [rows, C]; out: int16 [rows, C]. */
IVIT_API int ivit_layernorm_i8_i16x2(ivit_ctx*, const int8_t* x, int64_t rows, int C, const int32_t* bias_int,
                                     const ivit_dyadic_t* me, ivit_dyadic_t me2, int16_t* out, ivit_stream stream);

/* int8 -> int16 storage (values unchanged): the 8-bit output of PatchMerging's qact2 (swin_quant.py:347) entering the int16
 * residual stream of the next stage.  n % 16 == 0. */
IVIT_API int ivit_widen_i8_i16(ivit_ctx*, const int8_t* x, int64_t n, int16_t* out, ivit_stream stream);

/* Token average + QuantAct (swin_quant.py:554-555): out[b, c] = clamp8(RNE(RNE(sum_t x[b, t, c] / L) * m / 2^e)).
 * x: int8 [B, L, C], out: int8 [B, C]; C % 4 == 0. */
IVIT_API int ivit_avgpool_requant_i8(ivit_ctx*, const int8_t* x, int B, int L, int C, ivit_dyadic_t me, int8_t* out,
                                     ivit_stream stream);

/* ---- TVM-semantics compatibility mode -------------------------------------------------------
 * The integer row operators as the reference's DEPLOYMENT tree states them in Relay, for cross-checking against the
 * authors' deployed numerics: int32 wrapping arithmetic, truncating divisions, `(r >> 1) - x0` exponent, no clamp of the
 * sums.  They differ numerically from ivit_shiftmax / ivit_shiftgelu / ivit_layernorm above (which follow
 * models/quantization_utils/quant_modules.py) and are never used by the engines.  x0 = int(-1 / input_scale - 1)
 * (layers.py:357; for GELU input_scale * 1.702, layers.py:394), computed by the caller.
 *   ivit_tvm_softmax   replaces TVM_benchmark/models/layers.py:372-386 quantized_softmax (n = 16): int32 [rows, cols] -> int8
 *   ivit_tvm_gelu      replaces TVM_benchmark/models/layers.py:389-404 quantized_gelu (n = 23): int32 -> int32 (pre * sigmoid_int)
 *   ivit_tvm_layernorm replaces TVM_benchmark/models/layers.py:329-350 quantized_layernorm: int32 [rows, C] + bias_int[C] -> int32 */
IVIT_API int ivit_tvm_softmax(ivit_ctx*, const int32_t* x, int64_t rows, int cols, int32_t x0, int n, int8_t* out,
                              ivit_stream stream);
IVIT_API int ivit_tvm_gelu(ivit_ctx*, const int32_t* x, int64_t rows, int cols, int32_t x0, int n, int32_t* out,
                           ivit_stream stream);
IVIT_API int ivit_tvm_layernorm(ivit_ctx*, const int32_t* x, int64_t rows, int C, const int32_t* bias_int, int32_t* out,
                                ivit_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* IVIT_B200_H */
