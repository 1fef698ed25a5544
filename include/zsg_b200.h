/* zsg_b200.h — C ABI of libzsg_b200.so: the ZSGNet per-batch hot path on B200 (sm_100a).
 *
 * The reference (TheShadow29/zsgnet-pytorch) has no native code and no FFI; every entry
 * point below replaces a *library-op call site* of the reference hot path (SURVEY.md 2b /
 * section 8a).  Each declaration cites the reference lines it stands in for.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless said otherwise;
 *  - activations are NHWC float32 ("rows" = pixels, channels contiguous);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates
 *    nothing and keeps no global state; the caller owns every buffer;
 *  - return value: 0 on success, negative ZSG_E* code otherwise; zsg_last_error_string()
 *    gives the text of the last failure on the calling thread.
 */
#ifndef ZSG_B200_H
#define ZSG_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZSG_OK 0
#define ZSG_EINVAL (-1)   /* bad argument (shape, alignment, null pointer)          */
#define ZSG_ECUDA (-2)    /* a CUDA runtime error was raised by the launch          */
#define ZSG_EARCH (-3)    /* device is not sm_100 (tcgen05 path unavailable)        */

typedef void* zsg_stream_t;

const char* zsg_last_error_string(void);
int zsg_abi_version(void);
/* 1 if the current device can run the tcgen05 kernels (compute capability 10.x). */
int zsg_device_supported(void);

/* ----------------------------------------------------------------------------------------
 * Row table: one 16-byte entry per GEMM row (= output pixel of a conv, or input pixel of a
 * dgrad).  It makes the implicit-GEMM kernels geometry-agnostic: forward conv, data-gradient
 * (stride 1 or 2), and the six-level shared head (mdl.py:379-380) are the same kernel.
 *   base   element offset of this row's image plane (b,0,0,0) inside the input tensor
 *   y0,x0  input coordinate of filter tap (0,0)
 *   hin,win input plane size for this row (levels differ inside one launch)
 *   out    element offset of this row in the output tensor
 * -------------------------------------------------------------------------------------- */
typedef struct {
  int32_t base;
  int16_t y0, x0;
  int16_t hin, win;
  int32_t out;
} zsg_row_t;

/* conv implicit GEMM:  y[row.out + n] = epi( sum_{r,s,c} pro(x[row.base + ((y0+r*dil)/div*win + (x0+s*dil)/div)*cin + c]) * w[n][r][s][c] )
 * Replaces nn.Conv2d forward and its data-gradient on: torchvision resnet50 convs (mdl.py:149-156),
 * FPN convs (fpn_resnet.py:157-172), head convs (mdl.py:235-244, 379-380), LSTM input projection
 * (mdl.py:319).  Arithmetic: 3xTF32 on tcgen05 tensor cores (fp32-accurate), fp32 accumulate in TMEM. */
typedef struct {
  const float* x;          /* input activations                                             */
  const float* w;          /* [cout][r][s][cin] (K-major)                                   */
  const float* w_lo;       /* optional: if set, `w` holds the TF32-exact high parts and w_lo the fp32 remainders
                              (zsg_split_tf32) and the weight operand is fetched by TMA instead of gathered */
  float* y;                /* output                                                        */
  const zsg_row_t* rows;   /* [m]                                                           */
  int32_t m, cin, cout, r, s;
  int32_t in_div;          /* 1: forward / stride-1 dgrad.  2: stride-2 dgrad (tap valid iff even) */
  const float* in_scale;   /* optional per-input-channel affine applied on load (BatchNorm) */
  const float* in_shift;
  int32_t in_relu;         /* ReLU after the affine on load                                 */
  const float* bias;       /* optional [cout]                                               */
  int32_t out_relu;
  const float* out_mask;   /* optional, indexed like y: result = 0 where out_mask <= 0 (ReLU backward),
                              applied after the bias and before residual / accumulate            */
  const float* residual;   /* optional, indexed like y                                      */
  int32_t accumulate;      /* y += result instead of y = result                             */
  int32_t impl;            /* 0 = tcgen05 (product path), 1 = SIMT check kernel (tests only); bf16 operand path: 2 / 3 force
                              128- / 256-column tiles (tests; 0 chooses by layer size) */
  const float* x_lo;       /* optional (needs w_lo, no in_scale / in_relu): fp32 remainders of x after TF32 truncation
                              (zsg_split_act), same indexing as x.  The input operand then goes global -> shared by
                              cp.async with no register pass: the tensor core reads x itself as the TF32 high part */
  int32_t dil;             /* tap spacing: tap (r,s) reads (y0 + r*dil, x0 + s*dil); 0 means 1.  6 for the SSD conv6
                              (ssd_vgg.py:129) and its data gradient                                              */
  float* stats;            /* optional (plain outputs only: no bias / ReLU / mask / residual / accumulate): BatchNorm
                              statistics of y as a by-product of the epilogue.  [ceil(m/128)*4][2][cout] floats: for every
                              32-row group of the output, the column sums and the column sums of squares.
                              zsg_bn_stats_partials turns them into the sums zsg_bn_finalize reads, without
                              re-reading y (replaces zsg_bn_stats for the 53 convs that feed a BatchNorm)             */
  const uint16_t* x_bf16;  /* optional pair (both or none; cin % 8 == 0; no in_scale / in_relu): bfloat16 images of the input   */
  const uint16_t* w_bf16;  /* (indexed like x, zsg_cast_bf16) and of the weights [cout][r][s][cin].  Selects the bf16 operand  */
                           /* path of BASELINE configs 3-5 ("bf16 tensor-core convs"): one kind::f16 MMA per product, fp32      */
                           /* accumulation, fp32 epilogue and output; x / w / x_lo / w_lo are not read then (may be NULL)      */
  int32_t x_plain;         /* 1: the gather is the identity -- r = s = 1, in_div = 1 and row i reads x[i * cin .. (i+1) * cin) (every  */
                           /* 1x1 stride-1 conv of the path and its data gradient; the caller built the table, so it knows).  With */
                           /* operand images (x_lo or x_bf16) the A tiles are then fetched by TMA like the weights: no per-thread */
                           /* address work, no row-table reads on the producer side.  0: unknown (always correct)                 */
  uint16_t* y_bf16;        /* optional ("bf16 storage" of the bf16 engine's trunk): the output is stored as bfloat16 (round to      */
                           /* nearest even), indexed like y, INSTEAD of y (y may then be NULL).  Plain outputs only: no bias /     */
                           /* ReLU / mask / residual / accumulate, cout % 8 == 0, row offsets % 8 == 0; `stats` still come from    */
                           /* the fp32 accumulators                                                                                */
  const uint16_t* residual_bf16; /* optional, instead of `residual`: the same addend read from a bfloat16 tensor indexed like y    */
                           /* (bf16 storage: the shortcut gradient of a bottleneck block); cout % 4 == 0, row offsets % 4 == 0     */
  int32_t y_pitch;         /* > 0: rows[i].out == i * y_pitch for every row (the output is a plain [m, y_pitch] matrix, true for  */
                           /* every forward table of the path but the [B,A,5] scatter): the epilogue then computes the offsets     */
                           /* instead of reading the table.  0: unknown (always correct)                                           */
  const float* row_add;    /* optional pair: result[row][n] += row_add[row_add_idx[2*row] + n] + row_add[row_add_idx[2*row+1] + n]  */
  const int32_t* row_add_idx; /* after the bias, before the ReLU (the language and grid terms of the first head conv); cout % 4  */
                           /* == 0, offsets % 4 == 0                                                                              */
  float* y_lo;             /* optional, one of the two: the GEMM operand image of the OUTPUT, written by the epilogue next to y     */
  uint16_t* y_img_bf16;    /* (indexed like y): y_lo = TF32 remainders as zsg_split_act writes them, y_img_bf16 = the bfloat16 copy  */
                           /* as zsg_cast_bf16.  The consumer of y (next conv, weight / data gradient) then needs no image pass.   */
                           /* Plain [m, y_pitch] outputs with cout % 8 == 0 and y_pitch % 2 == 0 on the operand-image paths only   */
  const void* bnb_x;       /* optional group (all four or none): this launch is the data gradient that writes dy of a BatchNorm+ReLU  */
  const float* bnb_scale;  /* pair (mdl.py:149-156: bn1 / bn2 of a bottleneck) whose INPUT x (= output of the conv in front of it)   */
  const float* bnb_shift;  /* is indexed like y: float32 with a float32 y, bfloat16 with y_bf16.  The epilogue then also writes, per */
  float* bnb_partials;     /* 32-row group, sum dz and sum dz * x with dz = dy where x * scale + shift > 0 ([ceil(m/128)*4][2][cout],  */
                           /* the layout of `stats`): zsg_bn_stats_partials + zsg_bn_bwd_center_sums turn them into the sums          */
                           /* zsg_bn_bwd_apply reads, replacing zsg_bn_bwd_reduce (mask_mode 1) and its read of dy and x.  y is       */
                           /* stored unmasked as without it.  Plain outputs, y_pitch > 0, cout % 8 == 0, no `stats`                   */
} zsg_conv_params;
int zsg_conv_fwd(const zsg_conv_params* p, zsg_stream_t stream);
/* diagnostics only (tools/trace_conv.py): CTA 0 of the conv kernel writes clock stamps of its first `nblocks`
 * K blocks to buf[nblocks * 16] (device memory); buf = NULL switches tracing off. */
int zsg_debug_set_conv_trace(unsigned int* buf, int nblocks);

/* weight gradient: dw[n][r][s][c] += sum_rows dy[row.out + n] * pro(x[gather(row,r,s) + c]).
 * Replaces the cuDNN wgrad inside loss.backward() (utils.py:412) for every conv above and the
 * LSTM weight gradients.  dw must be zeroed by the caller (split-K partials are added atomically). */
typedef struct {
  const float* x;
  const float* dy;
  float* dw;               /* [cout][r][s][cin]                                              */
  const zsg_row_t* rows;   /* the forward conv's table                                       */
  int32_t m, cin, cout, r, s;
  const float* in_scale;
  const float* in_shift;
  int32_t in_relu;
  int32_t split_k;         /* 0 = choose automatically                                       */
  int32_t impl;            /* as in zsg_conv_params (bf16 path: 2 / 3 force 128- / 256-column tiles) */
  const float* x_lo;       /* optional pair (both or none; no in_scale / in_relu): TF32 remainders of x and dy      */
  const float* dy_lo;      /* (zsg_split_act); operands then go global -> shared by cp.async                         */
  int32_t dy_pitch;        /* > 0: rows[i].out == i * dy_pitch for every row (dy is a plain [m, dy_pitch] matrix, true for
                              every forward table of the path): dy / dy_lo are then fetched by TMA.  0: unknown        */
  int32_t dil;             /* tap spacing as in zsg_conv_params (0 means 1)                                             */
  const uint16_t* x_bf16;  /* optional pair (both or none; cin % 8 == 0, dy_pitch > 0 and % 8 == 0): bfloat16 images of x and */
  const uint16_t* dy_bf16; /* dy (zsg_cast_bf16); bf16 operand path as in zsg_conv_params, dw stays fp32 (atomic split-K)      */
} zsg_wgrad_params;
int zsg_conv_wgrad(const zsg_wgrad_params* p, zsg_stream_t stream);

/* w [cout][r][s][cin] -> wt [cin][r][s][cout] with the taps flipped: the dgrad of a conv is the
 * forward kernel applied to wt. */
int zsg_weight_transpose_flip(const float* w, float* wt, int cout, int rs_r, int rs_s, int cin, zsg_stream_t stream);
/* The same for n weight tensors in ONE launch (the backward needs ~64 of them per step and each alone is a
 * launch-latency-bound 8 us kernel): entry e transposes src_base[src .. ) -> dst_base[dst .. ); `begin` is the
 * running element count (begin[0] = 0, ascending), total = sum of cout*r*s*cin.  descs lives in device memory. */
typedef struct {
  int64_t src, dst, begin;
  int32_t cout, r, s, cin;
} zsg_wtf_desc;
int zsg_weight_transpose_flip_batched(const float* src_base, float* dst_base, const zsg_wtf_desc* descs, int n,
                                      int64_t total, zsg_stream_t stream);
/* The same for tables in which EVERY cin and cout is a multiple of 32 (all of the path's): 32 x 32 tiles through shared memory,
 * coalesced on both sides. */
int zsg_weight_transpose_flip_batched32(const float* src_base, float* dst_base, const zsg_wtf_desc* descs, int n,
                                        int64_t total, zsg_stream_t stream);
/* Operand preparation for the cp.async GEMM paths: z = relu?(x * scale[c] + shift[c]) (z may be NULL when there is
 * no affine / ReLU: the tensor is its own high part) and lo = z - trunc_tf32(z).  x is [rows, c], c % 4 == 0.
 * Replaces the on-the-fly split inside the conv kernels for every conv of the path (mdl.py / fpn_resnet.py). */
int zsg_split_act(const float* x, const float* scale, const float* shift, int relu, float* z, float* lo, int64_t rows,
                  int c, zsg_stream_t stream);
/* bf16 operand image for the bf16 GEMM path (configs 3-5): out = bf16_rn( relu?(x * scale[c] + shift[c]) ), x [rows, c],
 * c % 4 == 0; scale / shift may be NULL.  With rows = n / 4, c = 4 it is the plain cast of a flat array (weights). */
int zsg_cast_bf16(const float* x, const float* scale, const float* shift, int relu, uint16_t* out, int64_t rows, int c,
                  zsg_stream_t stream);
/* hi[i] = w[i] with the 13 low mantissa bits cleared (exactly representable in TF32), lo[i] = w[i] - hi[i]. */
int zsg_split_tf32(const float* w, float* hi, float* lo, int64_t n, zsg_stream_t stream);
/* row-wise channel padding copy: dst[n][0:cdst] = src[n][0:csrc] (zero fill / truncate). */
int zsg_pad_channels(const float* src, float* dst, int64_t n, int csrc, int cdst, zsg_stream_t stream);
/* NCHW image -> NHWC with 4 channels (4th = 0) for the stem (mdl.py:149). */
int zsg_nchw_to_nhwc4(const float* img, float* out, int b, int h, int w, zsg_stream_t stream);
/* bf16 operand path: NCHW image -> NHWC with 8 bfloat16 channels (r, g, b, 0 x 5): padded image and GEMM operand image in one. */
int zsg_nchw_to_nhwc8_bf16(const float* img, uint16_t* out, int b, int h, int w, zsg_stream_t stream);
/* column sums: out[c] (+)= sum_rows x[row*ld + c]   (bias gradients of FPN/head convs and LSTM). */
int zsg_colsum(const float* x, float* out, int64_t rows, int c, int ld, int accumulate, zsg_stream_t stream);
/* dst[i][0:cdst] = src[rows[i].out + 0:csrc] (zero padded): turns the packed [B,A,5] head gradient
 * (mdl.py:246-254 layout) back into level-major rows for the last head conv's backward. */
int zsg_gather_rows(const float* src, const zsg_row_t* rows, float* dst, int64_t m, int csrc, int cdst,
                    zsg_stream_t stream);

/* ---------------------------------- BatchNorm (training) ---------------------------------
 * torchvision resnet50's 53 BatchNorm2d in train mode (mdl.py:149-156, utils.py:395).      */
/* sums[0:c] = sum x, sums[c:2c] = sum x^2 (double; must be zeroed by the caller). */
int zsg_bn_stats(const float* x, double* sums, int64_t rows, int c, zsg_stream_t stream);
/* the same sums from the per-row-group partials a conv wrote (zsg_conv_params.stats): parts = ceil(m/128)*4. */
int zsg_bn_stats_partials(const float* partials, int64_t parts, int c, double* sums, zsg_stream_t stream);
/* sums[c .. 2c) <- invstd * (sums[c .. 2c) - mean * sums[0 .. c)): turns (sum dz, sum dz * x) from a data gradient's epilogue
 * (zsg_conv_params.bnb_partials, reduced by zsg_bn_stats_partials) into (sum dz, sum dz * xhat), what zsg_bn_bwd_apply reads. */
int zsg_bn_bwd_center_sums(double* sums, const float* mean, const float* invstd, int c, zsg_stream_t stream);
/* zsg_bn_stats_partials + zsg_bn_finalize in one launch: the last block of a channel group to finish (ticket counter)
 * finalizes it.  sums: 2c doubles, zeroed by the caller; tickets: 64 ints, zero before the first use, left zero. */
int zsg_bn_finalize_partials(const float* partials, int64_t parts, int64_t rows, int c, const float* gamma,
                             const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                             float* mean, float* invstd, float* scale, float* shift, double* sums, int* tickets,
                             zsg_stream_t stream);
/* mean/invstd/scale/shift from the sums; updates running stats (momentum, unbiased var). */
int zsg_bn_finalize(const double* sums, int64_t rows, int c, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                    float* scale, float* shift, zsg_stream_t stream);
/* eval mode: scale/shift from running stats. */
int zsg_bn_eval_affine(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                       float eps, int c, float* scale, float* shift, zsg_stream_t stream);
/* y = relu?( x*scale+shift  [+ r*rscale+rshift | + r] ) : the bottleneck tail (bn3 + shortcut + ReLU).
 * y_lo (optional): TF32 remainder image of y for the GEMMs that read it (see zsg_split_act). */
int zsg_bn_apply(const float* x, const float* scale, const float* shift, const float* r, const float* rscale,
                 const float* rshift, int relu, float* y, float* y_lo, int64_t rows, int c, zsg_stream_t stream);
/* zsg_bn_apply that also writes the bf16 image of y (y_bf16, required) instead of a TF32 remainder image. */
int zsg_bn_apply_bf16(const float* x, const float* scale, const float* shift, const float* r, const float* rscale,
                      const float* rshift, int relu, float* y, uint16_t* y_bf16, int64_t rows, int c, zsg_stream_t stream);
/* backward reduce: dz = dy * mask ; sums[0:c] = sum dz, sums[c:2c] = sum dz*xhat.
 * mask_mode 0: none; 1: relu mask from (x*scale+shift) > 0; 2: relu mask from act_out > 0 (dz is also
 * written to dz_out when non-null, for the shortcut path). */
int zsg_bn_bwd_reduce(const float* dy, const float* x, const float* mean, const float* invstd, const float* scale,
                      const float* shift, const float* act_out, int mask_mode, float* dz_out, double* sums,
                      int64_t rows, int c, zsg_stream_t stream);
/* dx = gamma*invstd*(dz - s1/rows - xhat*s2/rows); dgamma = s2, dbeta = s1 (written once).
 * dx_lo (optional): TF32 remainder image of dx for the data- and weight-gradient GEMMs that read it. */
int zsg_bn_bwd_apply(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                     const float* scale, const float* shift, const float* act_out, int mask_mode,
                     const double* sums, float* dx, float* dx_lo, float* dgamma, float* dbeta, int64_t rows, int c,
                     zsg_stream_t stream);

/* ---- bf16 storage: trunk activations of the bf16 engine (conv outputs x, BatchNorm+ReLU images, block outputs) live in HBM as
 * bfloat16 ONLY; the same kernels with bfloat16 loads / stores, arithmetic in fp32 (sums in fp64) as above.
 * Under torch.autocast(bfloat16) the reference stores exactly these tensors in bfloat16 (mdl.py:149-156 trunk). ---- */
/* out = bf16( relu?(x * scale[c] + shift[c]) ), x bfloat16 [rows, c] (zsg_cast_bf16 with a bfloat16 source). */
int zsg_act_b16(const uint16_t* x, const float* scale, const float* shift, int relu, uint16_t* out, int64_t rows, int c,
                zsg_stream_t stream);
/* zsg_bn_apply over bfloat16 tensors: y = bf16( relu?( x*scale+shift [+ r*rscale+rshift | + r] ) ). */
int zsg_bn_apply_b16(const uint16_t* x, const float* scale, const float* shift, const uint16_t* r, const float* rscale,
                     const float* rshift, int relu, uint16_t* y, int64_t rows, int c, zsg_stream_t stream);
/* zsg_bn_bwd_reduce with x / act_out in bfloat16; dy and dz_out are float32 or bfloat16 (flags). */
int zsg_bn_bwd_reduce_b16(const void* dy, int dy_is_b16, const uint16_t* x, const float* mean, const float* invstd,
                          const float* scale, const float* shift, const uint16_t* act_out, int mask_mode, void* dz_out,
                          int dz_is_b16, double* sums, int64_t rows, int c, zsg_stream_t stream);
/* zsg_bn_bwd_apply_bf16 with x / act_out in bfloat16; dy float32 or bfloat16; writes only the bfloat16 dx. */
int zsg_bn_bwd_apply_b16(const void* dy, int dy_is_b16, const uint16_t* x, const float* mean, const float* invstd,
                         const float* gamma, const float* scale, const float* shift, const uint16_t* act_out, int mask_mode,
                         const double* sums, uint16_t* dx_bf16, float* dgamma, float* dbeta, int64_t rows, int c,
                         zsg_stream_t stream);
/* zsg_maxpool_bn_relu_fwd with a bfloat16 source and a bfloat16 result. */
int zsg_maxpool_bn_relu_fwd_b16(const uint16_t* x, const float* scale, const float* shift, uint16_t* y, uint8_t* argmax,
                                int b, int h, int w, int c, int ho, int wo, zsg_stream_t stream);

/* zsg_bn_bwd_apply that writes the bf16 image of dx (dx_bf16, required) instead of a TF32 remainder image; dx may be NULL
 * (the fp32 gradient is not stored when only the two GEMMs that follow, which read the image, consume it). */
int zsg_bn_bwd_apply_bf16(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                          const float* scale, const float* shift, const float* act_out, int mask_mode,
                          const double* sums, float* dx, uint16_t* dx_bf16, float* dgamma, float* dbeta, int64_t rows,
                          int c, zsg_stream_t stream);

/* ------------------------------ pooling / resampling glue -------------------------------- */
/* stem: y = maxpool3x3/2,p1( relu(x*scale+shift) )  (mdl.py:150-152).  argmax (optional, same shape as y, one
 * byte per element) records the winning tap dy*3+dx (first maximum in scan order, as ATen does). */
int zsg_maxpool_bn_relu_fwd(const float* x, const float* scale, const float* shift, float* y, uint8_t* argmax, int b,
                            int h, int w, int c, int ho, int wo, zsg_stream_t stream);
/* da = gradient w.r.t. relu(bn(x)) (the max-pool input), gathered through the recorded argmax codes. */
int zsg_maxpool_bn_relu_bwd(const uint8_t* argmax, const float* dy, float* da, int b, int h, int w, int c, int ho,
                            int wo, zsg_stream_t stream);
/* dst += nearest_up(src) with host-computed index tables (fpn_resnet.py:161-162,166-167). */
int zsg_upsample_add(float* dst, const float* src, const int32_t* idx_y, const int32_t* idx_x, int b, int ho,
                     int wo, int hi, int wi, int c, zsg_stream_t stream);
/* dsrc += sum over the dst pixels that read each src pixel. */
int zsg_upsample_add_bwd(const float* ddst, float* dsrc, const int32_t* idx_y, const int32_t* idx_x, int b, int ho,
                         int wo, int hi, int wi, int c, zsg_stream_t stream);
/* global average pool (fpn_resnet.py:177) and its gradient (accumulating). */
int zsg_avgpool_fwd(const float* x, float* y, int b, int hw, int c, zsg_stream_t stream);
int zsg_avgpool_bwd(const float* dy, float* dx, int b, int hw, int c, zsg_stream_t stream);
/* SSD-VGG trunk glue (config 5; ssd_vgg.py).
 * Generic NHWC max-pool: window k, stride, symmetric pad; (ho, wo) chosen by the caller (floor or ceil mode:
 * nn.MaxPool2d(2,2), MaxPool2d(2,2,ceil_mode=True), MaxPool2d(3,1,1) at ssd_vgg.py:115-118,127).  argmax (one byte per
 * output element) records the winning tap dy*k+dx, first maximum in scan order as ATen does. */
int zsg_maxpool_fwd(const float* x, float* y, uint8_t* argmax, int b, int h, int w, int c, int k, int stride, int pad,
                    int ho, int wo, zsg_stream_t stream);
/* dx = gradient w.r.t. the pool input, gathered through the codes; mask (optional, indexed like dx: the pool's
 * input, a ReLU output): dx = 0 where mask <= 0, i.e. the ReLU backward of vgg[k-1] folded in. */
int zsg_maxpool_bwd(const uint8_t* argmax, const float* dy, const float* mask, float* dx, int b, int h, int w, int c,
                    int k, int stride, int pad, int ho, int wo, zsg_stream_t stream);
/* y[row] = x[row] / ||x[row]||_2 over the c channels (ssd_vgg.py:80: no eps, no learned scale); norm[row] is kept. */
int zsg_l2norm_fwd(const float* x, float* y, float* norm, int64_t rows, int c, zsg_stream_t stream);
/* dx (+)= dy / n - x * sum_c(dy * x) / n^3, zeroed where x <= 0 when mask_relu (x is the output of vgg[22]). */
int zsg_l2norm_bwd(const float* dy, const float* x, const float* norm, float* dx, int64_t rows, int c, int accumulate,
                   int mask_relu, zsg_stream_t stream);
/* dx = (accumulate? dx : 0) + dy * (x > 0) */
int zsg_relu_bwd(const float* dy, const float* x, float* dx, int64_t n, int accumulate, zsg_stream_t stream);
int zsg_axpy(const float* x, float* y, float a, int64_t n, zsg_stream_t stream); /* y += a*x */
/* x[i*stride + 0:width] *= (float)*scale, scale a DEVICE double: applies the incoming autograd gradient
 * of the scalar loss (utils.py:410-412) without a host synchronisation. */
int zsg_scale_dev(float* x, int64_t rows, int width, int64_t stride, const double* scale, zsg_stream_t stream);

/* ---- the first head conv without the concatenated tensor (mdl.py:69-104 tiles the language vector and a coordinate grid over
 * every feature map, concatenates [feat | lang | grid] (514 channels) and convolves it; mdl.py:235-244).  Linear in its three
 * parts: conv(W, fused) = conv(W_f, feat) + L[b, border class of the cell] + G[cell], see csrc/elementwise.cu.  The 514-channel
 * tensor, its gradient and the K = 520 contractions over them never exist (a-6 at 0 bytes). ---- */
/* dst[row][0:c] = src[row][0:c], row pitches src_ld / dst_ld: the channel slices W_f / W_l / W_g of the [256][9][514] weight and
 * of its gradient. */
int zsg_copy_cols(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t rows, int c, zsg_stream_t stream);
/* forward terms: v [b][n*9] = lang x W_l^T (from zsg_conv_fwd), wg [n*9][2], gridpatch [total_cells][9][2] (grid value of the
 * neighbour cell under each tap, 0 outside the level) -> lang_cls [b][16][n] (sum of v over the taps valid for a border
 * class), grid_term [total_cells][n].  zsg_conv_params.row_add adds both per output row. */
int zsg_head0_lang_grid_terms(const float* v, const float* wg, const float* gridpatch, float* lang_cls, float* grid_term,
                              int b, int total_cells, int n, zsg_stream_t stream);
/* backward sums over dh [sum_l b*cells_l][n] (row of (sample, cell) = cell_base[cell] + sample * cell_stride[cell]):
 * tap_sums [b][n*9] = sum over the cells where the tap is valid (-> d lang and dW_l by two small GEMMs), and
 * dwg[(n*9+t) * dwg_ld + g] = dW_g (written into the [.., 514] gradient at column 512).  scratch: b * 8 * 34 * n floats.
 * Partial sums are added in a fixed order: deterministic. */
int zsg_head0_backward_sums(const float* dh, const int32_t* cell_base, const int32_t* cell_stride, const int32_t* cell_cls,
                            const float* gridpatch, int b, int total_cells, int n, float* scratch, size_t scratch_floats,
                            float* tap_sums, float* dwg, int64_t dwg_ld, zsg_stream_t stream);

/* -------------------- language/grid tiling fusion (mdl.py:69-104) ------------------------
 * fused[b][cell][0:256]=feat, [256:512]=lang[b], [512]=grid_y, [513]=grid_x, [514:cpad]=0, for the six
 * levels packed level-major: level l occupies rows [b*cells_l + cell] after lvl_row_off[l].            */
int zsg_fuse_lang_grid(const float* feat, const float* lang, const float* grid_yx, float* fused, int b,
                       int total_cells, const int32_t* lvl_cells /*[6]*/, int nlvl, int cfeat, int clang, int cpad,
                       zsg_stream_t stream);
/* dfeat = dfused[..., 0:256] ; dlang[b][j] = sum_cells dfused[b][cell][256+j] */
int zsg_unfuse_lang_grid(const float* dfused, float* dfeat, float* dlang, int b, int total_cells,
                         const int32_t* lvl_cells, int nlvl, int cfeat, int clang, int cpad, zsg_stream_t stream);

/* ---------------------------------- bi-LSTM (mdl.py:296-336) ----------------------------- */
/* forward direction recurrence over t < len[b]; gx = qvec*W_ih^T precomputed ([B,T,512]).
 * Saves gate activations / cell / previous-hidden sequences for BPTT.  lang[b][0:128] = h after len[b]. */
int zsg_lstm_fwd_dir(const float* gx, const float* whh_t /*[128][512]*/, const float* b_ih, const float* b_hh,
                     const float* h0 /*[B,128] per sample*/, const float* c0, const int32_t* lens, int b, int t,
                     float* gates /*[B,T,512]*/, float* cs /*[B,T,128]*/, float* hprev /*[B,T,128]*/,
                     float* lang /*[B,256]*/, zsg_stream_t stream);
/* reverse direction: ONE step on token len-1 from (h0r,c0r) (mdl.py:326-328 read position len-1). */
int zsg_lstm_rev_step(const float* qvec /*[B,T,E]*/, const float* wih /*[512][E]*/, const float* whh /*[512][128]*/,
                      const float* b_ih, const float* b_hh, const float* h0, const float* c0, const int32_t* lens,
                      int b, int t, int e, float* xlast /*[B,E]*/, float* gates /*[B,512]*/, float* lang /*[B,256]*/,
                      zsg_stream_t stream);
/* BPTT of the forward direction: dG [B,T,512] (zero for t>=len). */
int zsg_lstm_bwd_dir(const float* dlang /*[B,256]*/, const float* whh /*[512][128]*/, const float* gates,
                     const float* cs, const float* c0, const int32_t* lens, int b, int t, float* dgates,
                     zsg_stream_t stream);
int zsg_lstm_rev_step_bwd(const float* dlang, const float* gates /*[B,512]*/, const float* c0, int b,
                          float* dgates /*[B,512]*/, zsg_stream_t stream);

/* -------------------- anchor match + focal/smooth-L1 loss (loss.py:43-143) ----------------
 * Fused forward + gradient.  IoU in float64 with the reference's exact op order (anchors.py:90-116),
 * strict > thr, first-index argmax (bit-exact `pos`/`top1`).  att/reg may be strided views of one packed
 * [B,A,5] buffer: *_stride are element strides per anchor.
 * losses[0..2] = loss, cls_ls, box_ls (double); d_att/d_reg are gradients of `loss` (lamb_reg applied).
 * NaN follows torch.max: a NaN IoU row / NaN scores select the first NaN index, never an out-of-range one; a NaN
 * loss is replaced by the constants of loss.py:128-133 and the gradients are zeroed.
 * workspace: zsg_match_loss_workspace_bytes(B) bytes, 16-byte aligned; carries the positives' counts from the match
 * to the loss pass and nothing between calls.  a <= 32768.
 * Two halves, callable separately (the match only needs the annotations, so a training step runs it on a side stream
 * under the forward pass) or as one call:
 *   zsg_match      anchors.py:153-165 + loss.py:76-87: pos [B,A] (uint8), top1 [B], counts into the workspace.  fp64 ALU work.
 *   zsg_loss_grad  loss.py:88-143 and its gradient: the HBM-bound pass, 40 B per (sample, anchor); partial sums in fixed
 *                  order (bit-reproducible losses).  Must follow a zsg_match on the same workspace.                */
size_t zsg_match_loss_workspace_bytes(int b);
int zsg_match(const float* annot, const double* anchors, int b, int a, double match_thr, int use_multi, int64_t* top1,
              uint8_t* pos, void* workspace, size_t ws_bytes, zsg_stream_t stream);
int zsg_loss_grad(const float* att, int64_t att_stride, const float* reg, int64_t reg_stride, const float* annot,
                  const double* anchors, const uint8_t* pos, int b, int a, float alpha, float gamma, double lamb_reg,
                  double* losses, float* d_att, int64_t d_att_stride, float* d_reg, int64_t d_reg_stride,
                  void* workspace, size_t ws_bytes, zsg_stream_t stream);
int zsg_match_loss(const float* att, int64_t att_stride, const float* reg, int64_t reg_stride, const float* annot,
                   const double* anchors, int b, int a, double match_thr, float alpha, float gamma, double lamb_reg,
                   int use_multi, double* losses, float* d_att, int64_t d_att_stride, float* d_reg,
                   int64_t d_reg_stride, int64_t* top1, uint8_t* pos, void* workspace, size_t ws_bytes,
                   zsg_stream_t stream);

/* -------------------------- evaluator (evaluator.py:48-117) ------------------------------
 * best_ids = argmax sigmoid(att) (first index); Acc/MaxPos flags from the decoded box of that / the
 * IoU-argmax anchor; pred_boxes in pixel x1y1x2y2 (double); metrics[0]=Acc, [1]=MaxPos (float).       */
size_t zsg_eval_workspace_bytes(int b);
/* metrics holds 2 + 2*B floats (the per-sample flags follow the two means); workspace as for zsg_match_loss
 * (zsg_eval_workspace_bytes(B), all zero before the first call, left zero).  One kernel launch. */
int zsg_eval(const float* att, int64_t att_stride, const float* reg, int64_t reg_stride, const float* annot,
             const double* anchors, const float* img_size /*[B,2] (h,w)*/, int b, int a, double iou_thr,
             int64_t* best_ids, float* pred_scores, double* pred_boxes, float* metrics, void* workspace,
             size_t ws_bytes, zsg_stream_t stream);

/* ---------------------- data path (dat_loader.py:98-146) -----------------
 * What a CPU worker of the reference does per sample after JPEG decoding, for a whole batch on the device:
 * `img.resize((300, 300))` (Pillow ImagingResample on 8-bit RGB: BICUBIC, the default since Pillow 7.0, or NEAREST,
 * the default of the pinned pillow 6.1), `pil2tensor(img).float().div_(255)` (dat_loader.py:136), and the word-vector
 * lookup (dat_loader.py:115).  Bit-identical to Pillow: fixed-point coefficient tables are built on the host exactly like
 * Resample.c precompute_coeffs (zsg_b200/gpu_data.py) and passed in `tables`.
 *   src      packed decoded images, uint8 [h][w][3] each, image i at src_off
 *   tables   int32: per image and output column xx: {xmin, xmax, k[hksize]} from hk_off, per output row yy: {ymin - y_first,
 *            ymax, k[vksize]} from vk_off (NEAREST: xtab[out_w] at hk_off, ytab[out_h] at vk_off)
 *   workspace  bytes for the horizontal pass: n_rows * out_w * 3 per image at tmp_off (unused for NEAREST)
 *   out      float32 [n_images][3][out_h][out_w] in [0, 1]                                                        */
typedef struct {
  int64_t src_off, tmp_off;
  int32_t h, w;
  int32_t y_first, n_rows;   /* source rows the vertical pass reads: [y_first, y_first + n_rows) (Pillow: ybox_first / ybox_last) */
  int32_t hk_off, vk_off;    /* offsets into `tables` (in int32) */
  int32_t hksize, vksize;
} zsg_resize_desc;
int zsg_resize_rgb8(const uint8_t* src, const zsg_resize_desc* descs /*device*/, const int32_t* tables /*device*/,
                    int n_images, int out_h, int out_w, int max_rows /*max n_rows over the batch*/, int nearest,
                    uint8_t* workspace, float* out, zsg_stream_t stream);
/* out[i][0:dim] = table[tokens[i]][0:dim], zeros where tokens[i] < 0; dim % 4 == 0. */
int zsg_embed_gather(const int32_t* tokens, const float* table, float* out, int64_t n_tokens, int dim,
                     zsg_stream_t stream);

/* ------------------------------ Adam (main_dist.py:50) ----------------------------------- */
int zsg_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
             float eps, int step, float grad_scale, zsg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ZSG_B200_H */
