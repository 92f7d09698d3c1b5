/*
 * ddsp_b200.h -- C ABI of the B200-native DDSP-SVC synthesizer forward path.
 *
 * The reference (tarepan/DDSP-SVC-official) has no FFI: the path sits behind the Python
 * `torch.nn.Module` classes `Sins` / `CombSub` / `CombSubFast` (ddsp/vocoder.py:372-550) and
 * the functions of ddsp/core.py they call.  This header is the boundary a binding (ctypes,
 * a TORCH_LIBRARY op, cffi ...) talks to; every entry point names the reference code it
 * replaces.  See INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - plain C: pointers, sizes, strides (in ELEMENTS, not bytes); no C++/torch types.
 *   - every pointer is a DEVICE pointer on the current CUDA device unless stated otherwise.
 *   - every function returns 0 on success or a negative ddsp_b200_status; nothing throws.
 *   - nothing here allocates or frees device memory, and nothing synchronises: work is
 *     enqueued on `stream` (a cudaStream_t passed as void*) and the call returns.
 *   - re-entrant and thread-safe (no global mutable state besides immutable tables that are
 *     uploaded lazily once per device under a lock).
 *   - shapes: B clips, F frames per clip, hop = block_size samples per frame, T = F*hop.
 */
#ifndef DDSP_B200_H_
#define DDSP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDSP_B200_ABI_VERSION 1

typedef enum ddsp_b200_status {
    DDSP_B200_OK = 0,
    DDSP_B200_ERR_INVALID_ARGUMENT = -1,  /* null pointer, non-positive size, bad stride        */
    DDSP_B200_ERR_UNSUPPORTED = -2,       /* e.g. hop != 512 or n_mag not in {256,512} on a fused path */
    DDSP_B200_ERR_WORKSPACE = -3,         /* workspace too small                                  */
    DDSP_B200_ERR_CUDA = -4,              /* a CUDA runtime call failed (see ddsp_b200_last_cuda_error) */
    DDSP_B200_ERR_BATCH_MISMATCH = -5,    /* core.py:212-213 ValueError                            */
    DDSP_B200_ERR_CAPTURE = -6            /* first call on a device (table setup) issued under stream capture */
} ddsp_b200_status;

/* filter window modes of ddsp/core.py:306-328 */
#define DDSP_B200_WINDOW_NONE 0    /* hann_window=False           (core.py:324-326) */
#define DDSP_B200_WINDOW_HANN 1    /* hann_window=True            (core.py:242-289) */
#define DDSP_B200_WINDOW_DYNAMIC 2 /* half_width_frames != None   (core.py:292-303) */

/* magnitude encodings accepted by ddsp_b200_frequency_filter */
#define DDSP_B200_MAG_REAL 0        /* magnitudes are real >= 0: complex(m, 0)                  */
#define DDSP_B200_MAG_EXP 1         /* control c -> exp(c) * scale (vocoder.py:399,475,522-523)  */
#define DDSP_B200_MAG_ALLPASS_TANH 2 /* control c -> exp(j*cumsum(pi*tanh(c))) (vocoder.py:398,415,521,540) */
#define DDSP_B200_MAG_COMPLEX 3     /* interleaved complex64 magnitudes                           */

int ddsp_b200_version(void);                 /* DDSP_B200_ABI_VERSION of the loaded library */
const char *ddsp_b200_strerror(int status);  /* static string                                */
int ddsp_b200_last_cuda_error(void);         /* cudaError_t of the last failing call on this thread */

/* ------------------------------------------------------------------------------------------
 * Standalone ops of ddsp/core.py (used by callers outside the fused path and by the tests).
 * ---------------------------------------------------------------------------------------- */

/* upsample(signal, factor)                                        ddsp/core.py:7-21
 * x: (B,F,C) with element strides (sB,sF,sC) -> y: (B,F*factor,C) contiguous.
 * Bit-identical to torch's upsample_linear1d(align_corners=True) + hold-last. */
int ddsp_b200_upsample(const float *x, int64_t sB, int64_t sF, int64_t sC, int B, int F, int C,
                       int factor, float *y, void *stream);

/* fo_to_rot(fo, sr, initial_phase, precise)                       ddsp/core.py:31-51
 * fo: (B,T) contiguous -> rot: (B,T) contiguous, wrapped to [-0.5,0.5] (round-half-even).
 * initial_phase: (B,) radians or NULL.  precise!=0: fp64 accumulation (inference),
 * precise==0: fp32 accumulation (core.py:40).
 * workspace: ddsp_b200_fo_to_rot_workspace_bytes(B,T) bytes. */
size_t ddsp_b200_fo_to_rot_workspace_bytes(int B, int64_t T);
int ddsp_b200_fo_to_rot(const float *fo, int B, int64_t T, double sr, const float *initial_phase,
                        int precise, float *rot, void *workspace, size_t workspace_bytes,
                        void *stream);

/* remove_above_fmax(amplitudes, pitch, fmax, level_start)        ddsp/core.py:24-28
 * amplitudes (B,F,K) strides (aB,aF,1); pitch (B,F) strides (pB,pF); out (B,F,K) contiguous.
 * out = amp * ((pitch*k < fmax) + 1e-7), k = level_start..level_start+K-1, fp32, bit-exact. */
int ddsp_b200_remove_above_fmax(const float *amplitudes, int64_t aB, int64_t aF, const float *pitch,
                                int64_t pB, int64_t pF, float fmax, int level_start, int B, int F,
                                int K, float *out, void *stream);

/* frequency_filter(audio, magnitudes, hann_window, half_width_frames)   ddsp/core.py:331-336
 * (= _frequency_impulse_response :306-328 + _fft_convolve :185-239).
 * audio (B,T) contiguous, T = F*hop, hop must be 512; mags (B,F,n_mag) strides (mB,mF,1),
 * n_mag in {256,512}; encoding per DDSP_B200_MAG_*; `mag_scale` multiplies MAG_EXP results;
 * f0_frames (B,F) strides (fB,fF) only for DDSP_B200_WINDOW_DYNAMIC
 * (half_width_frames = 1.5*sr/(f0+1e-3), vocoder.py:542); out (B,T) contiguous, must not alias
 * audio.  `accumulate` != 0 adds into out instead of overwriting it. */
size_t ddsp_b200_frequency_filter_workspace_bytes(int B, int F, int n_mag);
int ddsp_b200_frequency_filter(const float *audio, const float *mags, int64_t mB, int64_t mF,
                               int n_mag, int mag_encoding, float mag_scale, int window_mode,
                               const float *f0_frames, int64_t fB, int64_t fF, double sr, int B,
                               int F, int hop, float *out, int accumulate, void *workspace,
                               size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Stage A of every synthesizer forward: f0 upsample + phase accumulation.
 *   Sins.forward        vocoder.py:391-393      CombSubFast.forward  vocoder.py:449-451
 *   CombSub.forward     vocoder.py:515-517      (upsample core.py:7-21, fo_to_rot core.py:31-51)
 * f0_frames (B,F) strides (fB,fF).  Outputs: phase_frames (B,F) contiguous
 * = fl32(2*pi)*rot[:, ::hop]; prefix (B,F) fp64 = sum of the upsampled fp32 f0 over all samples
 * before each frame (consumed by stage B); phase_full (B,T) or NULL (Sins: 2*pi*rot at sample
 * rate, vocoder.py:392).  hop must be 512.
 * `precise` mirrors fo_to_rot's flag (the forwards pass `infer`) and is accepted for signature compatibility only:
 * this fused stage ALWAYS accumulates in fp64.  The reference's precise=False (training, core.py:40) is an fp32
 * `cumsum` whose rounding depends on the scan order of the backend (torch CPU and CUDA already differ) and drifts by
 * whole rotations on long clips; the fp64 result is the value both approximate.  Callers that need the fp32
 * arithmetic itself use ddsp_b200_fo_to_rot(precise = 0).  `initial_phase` (radians, per clip or NULL) enters the
 * prefix here; the stage-B entry points take the prefix and ignore their own `initial_phase` argument.
 * ---------------------------------------------------------------------------------------- */
int ddsp_b200_phase(const float *f0_frames, int64_t fB, int64_t fF, int B, int F, int hop, double sr,
                    const float *initial_phase, int precise, float *phase_frames, double *prefix,
                    float *phase_full, void *stream);

/* ------------------------------------------------------------------------------------------
 * Stage B of CombSubFast.forward                               vocoder.py:455-492
 * ctrl views harmonic_magnitude / harmonic_phase / noise_magnitude: three pointers into
 * (B,F,513) views sharing strides (cB,cF,1) -- exactly what split_to_dict emits
 * (unit2control.py:10-20).  noise_u: (B,T) uniform [0,1) tensor standing in for
 * torch.rand_like (vocoder.py:461), or NULL to draw it in-kernel from (seed, clip, sample).
 * window: the module's `window` buffer sqrt(hann_window(1024)) (vocoder.py:434; 1024 floats) or
 * NULL for the exactly computed sin(pi*i/1024).  prefix: from ddsp_b200_phase.
 * signal: (B,T) contiguous.
 * ---------------------------------------------------------------------------------------- */
int ddsp_b200_combsubfast(const float *harmonic_magnitude, const float *harmonic_phase,
                          const float *noise_magnitude, int64_t cB, int64_t cF,
                          const float *f0_frames, int64_t fB, int64_t fF, const double *prefix,
                          const float *initial_phase, const float *noise_u, uint64_t seed,
                          const float *window, int B, int F, int hop, double sr, float *signal,
                          void *stream);

/* ------------------------------------------------------------------------------------------
 * Streaming forms of the two calls above: a block of frames that CONTINUES a
 * stream instead of starting at phase 0 -- what gui.py:373-388 emulates by re-synthesising its
 * whole window every block and splicing with SOLA (gui.py:408-426).
 *   ddsp_b200_phase_stream: `carry` points at the fp64 prefix the stream reached at this block's
 *     first frame (element b at carry[b*carry_stride]; typically a column of the previous block's
 *     `prefix` output, which must not be the same buffer as this call's) or NULL at stream start
 *     (then initial_phase applies as in core.py:44-45).
 *   ddsp_b200_combsubfast_stream: `hop_offset` = stream index of this block's first hop; it only
 *     shifts the in-kernel noise stream (ignored with an injected noise_u) so that a hop that is
 *     synthesised again as context of the next block gets the same noise.  `seed_device`: optional
 *     (NULL) device pointer to one uint64 that is added to `seed` when the kernel starts, so a call
 *     captured in a CUDA graph draws fresh noise on every replay (the caller bumps the counter
 *     inside the graph).
 * A stream of blocks, each overlapping its predecessor by 3 frames and keeping hops 1..F-3 of
 * every block, equals one call over the concatenated frames to the last fp32 ulp or two (same
 * phase, excitation and noise; the pair-packed inverse FFT rounds differently when a block starts
 * on an odd frame) -- tests/test_gpu_stream.py; the host side of that bookkeeping is
 * ddsp_b200.streaming.
 * ---------------------------------------------------------------------------------------- */
int ddsp_b200_phase_stream(const float *f0_frames, int64_t fB, int64_t fF, int B, int F, int hop,
                           double sr, const float *initial_phase, const double *carry,
                           int64_t carry_stride, float *phase_frames, double *prefix, void *stream);
int ddsp_b200_combsubfast_stream(const float *harmonic_magnitude, const float *harmonic_phase,
                                 const float *noise_magnitude, int64_t cB, int64_t cF,
                                 const float *f0_frames, int64_t fB, int64_t fF, const double *prefix,
                                 const float *noise_u, uint64_t seed, const uint64_t *seed_device,
                                 int64_t hop_offset, const float *window, int B, int F, int hop,
                                 double sr, float *signal, void *stream);

/* ------------------------------------------------------------------------------------------
 * Gradient of stage B of CombSubFast.forward with respect to the three control tensors -- what
 * autograd derives for vocoder.py:455-492 when the module is trained (solver.py:111,113).
 * Inputs are those of the forward call (same noise_u or seed, same window, same prefix) plus
 * grad_signal (B,T) contiguous = dL/dsignal.  Outputs: three (B,F,513) views sharing strides
 * (gB,gF,1), fully overwritten.  f0 / phase receive no gradient.
 * ---------------------------------------------------------------------------------------- */
int ddsp_b200_combsubfast_backward(const float *harmonic_magnitude, const float *harmonic_phase,
                                   const float *noise_magnitude, int64_t cB, int64_t cF,
                                   const float *f0_frames, int64_t fB, int64_t fF,
                                   const double *prefix, const float *noise_u, uint64_t seed,
                                   const float *window, const float *grad_signal, int B, int F,
                                   int hop, double sr, float *grad_harmonic_magnitude,
                                   float *grad_harmonic_phase, float *grad_noise_magnitude,
                                   int64_t gB, int64_t gF, void *stream);

/* ------------------------------------------------------------------------------------------
 * Stage B of CombSub.forward (old)                             vocoder.py:521-550
 * ctrl views group_delay (n_mag_allpass), harmonic_magnitude (n_mag_harmonic),
 * noise_magnitude (n_mag_noise) with common strides (cB,cF,1).  Outputs signal, harmonic,
 * noise: (B,T) contiguous each.  workspace: ddsp_b200_combsub_workspace_bytes.
 * ---------------------------------------------------------------------------------------- */
size_t ddsp_b200_combsub_workspace_bytes(int B, int F, int n_mag_allpass, int n_mag_harmonic,
                                         int n_mag_noise);
int ddsp_b200_combsub(const float *group_delay, int n_mag_allpass, const float *harmonic_magnitude,
                      int n_mag_harmonic, const float *noise_magnitude, int n_mag_noise, int64_t cB,
                      int64_t cF, const float *f0_frames, int64_t fB, int64_t fF,
                      const double *prefix, const float *initial_phase, const float *noise_u,
                      uint64_t seed, int B, int F, int hop, double sr, float *signal,
                      float *harmonic, float *noise, void *workspace, size_t workspace_bytes,
                      void *stream);

/* ------------------------------------------------------------------------------------------
 * Stage B of Sins.forward                                      vocoder.py:397-423
 * ctrl views amplitudes (n_harmonics), group_delay (n_mag_allpass), noise_magnitude
 * (n_mag_noise) with common strides (cB,cF,1).  phase_full: (B,T) from ddsp_b200_phase.
 * ---------------------------------------------------------------------------------------- */
size_t ddsp_b200_sins_workspace_bytes(int B, int F, int n_harmonics, int n_mag_allpass,
                                      int n_mag_noise);
int ddsp_b200_sins(const float *amplitudes, int n_harmonics, const float *group_delay,
                   int n_mag_allpass, const float *noise_magnitude, int n_mag_noise, int64_t cB,
                   int64_t cF, const float *f0_frames, int64_t fB, int64_t fF,
                   const float *phase_full, const float *noise_u, uint64_t seed, int B, int F,
                   int hop, double sr, float *signal, float *harmonic, float *noise,
                   void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Caller-side epilogue, fused:  signal *= upsample(mask_frames, hop)   main.py:116,159  gui.py:112,127
 * signal (B,T) contiguous, modified in place; mask_frames (B,F) strides (mB,mF).  Bit-identical to
 * `signal * upsample(mask)[...,0]` of the reference without materialising the (B,T) mask.
 * ---------------------------------------------------------------------------------------- */
int ddsp_b200_apply_frame_mask(float *signal, const float *mask_frames, int64_t mB, int64_t mF, int B,
                               int F, int hop, void *stream);

/* The whole silence-mask epilogue of the callers in one in-place pass            main.py:112-116,159 / gui.py:108-112,127
 *   mask = (volume > threshold) -> padded with its edge values by 4 frames -> 9-frame running maximum
 *        -> upsample(mask, block_size) -> signal *= mask
 * volume_frames: (B,F) view (element strides vB, vF), threshold = 10^(dB/20) compared in double as numpy does;
 * signal (B, F*hop) contiguous, 16-byte aligned, modified in place. */
int ddsp_b200_apply_volume_mask(float *signal, const float *volume_frames, int64_t vB, int64_t vF,
                                double threshold, int B, int F, int hop, void *stream);

/* Streaming forms of the frequency_filter models (ddsp/vocoder.py:381-423, :504-550): the same computation on a window
 * of frames that continues a stream -- `prefix` / `phase_full` come from ddsp_b200_phase_stream(_full) with the
 * carried fp64 prefix, and `hop_offset` is the stream index of the window's first hop, so that the in-kernel noise of
 * a re-synthesised hop repeats.  ddsp_b200.streaming.FilterModelStream re-synthesises the few frames whose output the
 * (L-1)-tap filters had not finished (2 + 3 frames for CombSub, 1 + 2 for Sins) and emits only finished hops. */
int ddsp_b200_phase_stream_full(const float *f0_frames, int64_t fB, int64_t fF, int B, int F, int hop, double sr,
                                const float *initial_phase, const double *carry, int64_t carry_stride,
                                float *phase_frames, double *prefix, float *phase_full, void *stream);
int ddsp_b200_combsub_stream(const float *group_delay, int n_mag_allpass, const float *harmonic_magnitude,
                             int n_mag_harmonic, const float *noise_magnitude, int n_mag_noise, int64_t cB,
                             int64_t cF, const float *f0_frames, int64_t fB, int64_t fF, const double *prefix,
                             const float *noise_u, uint64_t seed, int64_t hop_offset, int B, int F, int hop,
                             double sr, float *signal, float *harmonic, float *noise, void *workspace,
                             size_t workspace_bytes, void *stream);
int ddsp_b200_sins_stream(const float *amplitudes, int n_harmonics, const float *group_delay, int n_mag_allpass,
                          const float *noise_magnitude, int n_mag_noise, int64_t cB, int64_t cF,
                          const float *f0_frames, int64_t fB, int64_t fF, const float *phase_full,
                          const float *noise_u, uint64_t seed, int64_t hop_offset, int B, int F, int hop, double sr,
                          float *signal, float *harmonic, float *noise, void *workspace, size_t workspace_bytes,
                          void *stream);

/* ------------------------------------------------------------------------------------------
 * Control network (ddsp/unit2control.py, ddsp/pcmer.py) -- fused elementwise stages between the
 * library GEMMs.  Not part of the synthesizer path; see DESIGN.md section 7.
 *
 * FAVOR+ softmax-kernel feature map                               pcmer.py:124-160
 *   dash (B,N,H,M) contiguous = (64^-0.25 * x) @ projection^T;  x (B,N,H,64) contiguous (to_q / to_k
 *   output before the head split);  out (B,H,N,M) contiguous:
 *   query: M^-0.5 * (exp(dash - |x|^2/16 - max_j dash) + eps);  key: M^-0.5 * exp(dash - |x|^2/16 + eps).
 *   M <= 384.
 * GLU -> depthwise Conv1d(k=31, padding 'same') -> SiLU          pcmer.py:53-55, channels last
 *   u (B,T,2C) contiguous, weight (C,31), bias (C), out (B,T,C) contiguous.
 * ---------------------------------------------------------------------------------------- */
int ddsp_b200_performer_features(const float *dash, const float *x, int B, int N, int H, int M,
                                 int is_query, float eps, float *out, void *stream);
/* Same feature map with the projection fused (dash never stored): x (B,N,H,64) contiguous,
 * projection (M,64) contiguous, 256 <= M <= 288 (the network's M = int(64 ln 64) = 266), out (B,H,N,M).  x_bias (H*64) / u_bias (2C): optional (NULL)
 * bias of the Linear that produced x / u, added on load so that GEMM can run without a bias epilogue. */
int ddsp_b200_performer_project_features(const float *x, const float *x_bias, const float *projection,
                                         int B, int N, int H, int M, int is_query, float eps,
                                         float *out, void *stream);
int ddsp_b200_glu_dwconv_silu(const float *u, const float *u_bias, const float *weight,
                              const float *bias, int B, int T, int C, float *out, void *stream);

/* Whole non-causal Performer attention for streaming-sized blocks (after the q/k/v projections):
 * feature maps, k_sum, context, normalised output, head-merged.  pcmer.py:69-78,124-160,191-251
 * q, k, v: (B,N,H*64) views with `row_stride` elements between consecutive frames (H*64 when
 * contiguous; 3*H*64 for the three slices of one merged q|k|v projection) and B*N*row_stride per
 * clip; q_bias / k_bias / v_bias (H*64) optional (NULL): biases of the
 * producing Linears, added on load; projection (M,64), M <= 288; out (B,N,H*64) contiguous.
 * Blocks of up to 16 frames run as one kernel (one CTA per clip and head); longer ones as three
 * kernels over 8-frame tiles (partial contexts -> fixed-order sum -> outputs), which need the
 * workspace. */
size_t ddsp_b200_performer_attention_workspace_bytes(int B, int N, int H);
int ddsp_b200_performer_attention(const float *q, const float *k, const float *v, int64_t row_stride,
                                  const float *q_bias, const float *k_bias, const float *v_bias,
                                  const float *projection, int B, int N, int H, int M, float eps,
                                  float *out, void *workspace, size_t workspace_bytes, void *stream);

/* Input embedding sum of Unit2Control.forward                     unit2control.py:80-95
 *   out[b,n,c] = x[b,n,c] + f0_embed(log(1 + f0/700)) + phase_embed(phase/pi) + volume_embed(volume) + spk[b,c]
 * x: (B,N,C) view with element strides (xB,xN,xC); f0 / phase / volume: (B,N) views with strides;
 * w* / b*: weight (C) and bias (C) of the three Linear(1,C); spk: speaker row(s) (1,C) [sB = 0] or
 * (B,C) [sB = C]; out (B,N,C) contiguous. */
int ddsp_b200_embed_sum(const float *x, int64_t xB, int64_t xN, int64_t xC, const float *f0, int64_t fB,
                        int64_t fN, const float *phase, int64_t pB, int64_t pN, const float *volume,
                        int64_t vB, int64_t vN, const float *w_f0, const float *b_f0,
                        const float *w_phase, const float *b_phase, const float *w_volume,
                        const float *b_volume, const float *spk, int64_t sB, int B, int N, int C,
                        float *out, void *stream);

/* Unit pre-net on the tensor cores                                 unit2control.py:38-45
 *   Transpose - Conv1d(n_unit,256,3,'same') - GroupNorm(4,256) - LeakyReLU - Conv1d(256,256,3,'same') - Transpose
 * in channels-last layout without transposes: `ddsp_b200_pad_frames` copies the (B,N,C) frames (element strides xB, xN; unit
 * channel stride) into a contiguous (B,N+2,C) buffer with one zero frame in front of and behind every clip.  The window of
 * frame n is then the 3*C contiguous floats starting at padded frame n, so each convolution is ONE call of
 * `ddsp_b200_linear_tf32x3_ex` with A = the padded buffer, lda = C, K = 3*C, M = B*(N+2) - 2 and the weight laid out as
 * W'[o][t*C + c] = W[o][c][t]; the output of the first convolution is written one row further down into the next padded
 * buffer (its two garbage rows per clip boundary land on the pad frames).
 * `ddsp_b200_groupnorm_leaky` normalises such a padded (B,N+2,C) buffer in place over the N real frames of every clip
 * (GroupNorm statistics per clip and group in fp64: biased variance, eps inside the root), applies gamma / beta and
 * LeakyReLU(slope) and writes zeros to the pad frames.  sums: scratch of 2*groups*B doubles.
 * `ddsp_b200_embed_sum_ln` (C = 256) is `ddsp_b200_embed_sum` (unit2control.py:80-95) that also returns the first
 * LayerNorm of PCmer (pcmer.py:25) of every finished row: out, out_ln (B,N,256) contiguous. */
int ddsp_b200_pad_frames(const float *x, int64_t xB, int64_t xN, int B, int N, int C, float *out, void *stream);
int ddsp_b200_groupnorm_leaky(float *hp, const float *gamma, const float *beta, float eps, float slope, int groups,
                              int B, int N, int C, double *sums, void *stream);
int ddsp_b200_embed_sum_ln(const float *x, int64_t xB, int64_t xN, const float *f0, int64_t fB, int64_t fN,
                           const float *phase, int64_t pB, int64_t pN, const float *volume, int64_t vB, int64_t vN,
                           const float *w_f0, const float *b_f0, const float *w_phase, const float *b_phase,
                           const float *w_volume, const float *b_volume, const float *spk, int64_t sB,
                           const float *ln_gamma, const float *ln_beta, float ln_eps, int B, int N, int C,
                           float *out, float *out_ln, void *stream);

/* Linear layer on the tensor cores, fp32-faithful ("3xTF32": each operand split into two TF32 terms, three
 * tcgen05.mma per k-step accumulate hi*hi + hi*lo + lo*hi in fp32 TMEM; csrc/gemm_tc.cuh):
 *   C[m,n] = sum_k A[m,k] * W[n,k] (+ bias[n]) (+ residual[m,n])
 * Replaces the nn.Linear / 1x1 Conv1d calls of the control network (ddsp/unit2control.py:56-62,
 * ddsp/pcmer.py:41-63, :191-251).  A: (M,K) row stride lda; W: (N,K) = nn.Linear.weight, row stride ldw;
 * bias (N) and residual (M,N; row stride ldr) optional (NULL); C: (M,N) row stride ldc; may alias residual.
 * A and W must be 16-byte aligned with lda, ldw multiples of 4 (TMA); C / residual of any 4-byte alignment
 * (128-bit stores are used when ldc, ldr are multiples of 4 and the bases 16-byte aligned). */
int ddsp_b200_linear_tf32x3(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias,
                            const float *residual, int64_t ldr, float *C, int64_t ldc, int M, int N,
                            int K, void *stream);

/* Second generation of the same GEMM (csrc/gemm_attn.cuh): rows leave through TMA stores (coalesced), the weight may
 * arrive already split into its two TF32 terms (W = hi, W_lo = W - hi; NULL: split in shared memory), and
 * LayerNorm(C) over the N <= 256 columns (torch.nn.LayerNorm: biased variance, ln_eps inside the sqrt;
 * ddsp/pcmer.py:25, :44, ddsp/unit2control.py:58) can be written as a second output C_ln while the finished row
 * is still in tensor memory.  ln_gamma / ln_beta / C_ln NULL: no LayerNorm.  C, C_ln, residual: 16-byte aligned
 * with row strides that are multiples of 4. */
int ddsp_b200_linear_tf32x3_ex(const float *A, int64_t lda, const float *W, const float *W_lo, int64_t ldw,
                               const float *bias, const float *residual, int64_t ldr, float *C, int64_t ldc,
                               const float *ln_gamma, const float *ln_beta, float ln_eps, float *C_ln,
                               int64_t ldc_ln, int M, int N, int K, void *stream);

/* First pointwise convolution of the conformer module with the GLU fused into the GEMM epilogue  (ddsp/pcmer.py:52-53):
 *   C[m, 128 t + c] = (y[m, 256 t + c] + bias[256 t + c]) * sigmoid(y[m, 256 t + 128 + c] + bias[256 t + 128 + c]),
 *   y = A (M,K) * W^T, W (N,K) with N a multiple of 256 and its rows (and `bias`) interleaved per 256-column tile:
 *   rows 256 t .. 256 t + 127 = value channels 128 t .. 128 t + 127, rows 256 t + 128 .. 256 t + 255 = their gate
 *   channels (N/2 + 128 t ...).  C: (M, N/2), row stride ldc (multiple of 4, 16-byte aligned base). */
int ddsp_b200_linear_glu(const float *A, int64_t lda, const float *W, const float *W_lo, int64_t ldw,
                         const float *bias, float *C, int64_t ldc, int M, int N, int K, void *stream);

/* Depthwise Conv1d(k=31, 'same') -> SiLU on an already gated channels-last tensor g (B,T,C) -> out (B,T,C)
 * (ddsp/pcmer.py:54-55; the second half of ddsp_b200_glu_dwconv_silu). */
int ddsp_b200_dwconv_silu(const float *g, const float *w, const float *bias, int B, int T, int C, float *out,
                          void *stream);

/* Performer (FAVOR+) self-attention of one PCmer layer as tensor-core GEMMs     ddsp/pcmer.py:69-78,124-160,191-251
 * (non-causal; dim_head 64, 266 random features padded to 272; Z = B * heads; Fp = frames rounded up to a multiple of 4)
 *
 * ddsp_b200_qkv_heads:      [q | k | v] = A (B*F, K) * [W_q; W_k; W_v]^T + bias in one GEMM; q, k stored head-major
 *                           (B,H,F,64); v stored transposed into rows 0..63 of vt (B,H,80,Fp) -- the caller presets
 *                           row 64 to ones and rows 65..79 / columns >= F to zero once.  vt_lo (optional, same layout,
 *                           zero-initialised once) receives v^T minus its TF32 truncation: the pre-split low term
 *                           ddsp_b200_favor_context takes for its first operand.
 * ddsp_b200_favor_features: x (Z,F,64) -> softmax-kernel features of dash = x * proj_scaled^T with
 *                           proj_scaled = 64^-0.25 * projection_matrix (266,64):
 *                             query: out (Z,F,272)   = 266^-0.5 (exp(dash - |x|^2/16 - max_j dash) + eps), pad columns 0
 *                             key:   out (Z,272,Fp)  = 266^-0.5  exp(dash - |x|^2/16 + eps), transposed; pad rows /
 *                                    columns are never written (zero-initialise the buffer once)
 * ddsp_b200_favor_context:  ctxT[z] (80,272) = vt[z] (80,Fp) * kt[z] (272,Fp)^T   (row 64 = sum over frames of k');
 *                           ctxT_lo (optional, same shape) receives ctxT minus its TF32 truncation, which
 *                           ddsp_b200_favor_output then takes as the pre-split low term of its second operand
 * ddsp_b200_favor_output:   out (B,F,H*64): out[b,f,h*64+e] = (q'[z,f,:] . ctxT[z,e,:]) / (q'[z,f,:] . ctxT[z,64,:] + 1e-8) */
int ddsp_b200_qkv_heads(const float *A, int64_t lda, const float *W, const float *W_lo, int64_t ldw,
                        const float *bias, float *q, float *k, float *vt, float *vt_lo, int B, int F, int Fp, int H,
                        int K, void *stream);
int ddsp_b200_favor_features(const float *x, const float *proj_scaled, int n_features, int is_query, float eps,
                             float *out, int Z, int F, int Fp, void *stream);
int ddsp_b200_favor_context(const float *vt, const float *vt_lo, const float *kt, float *ctxT, float *ctxT_lo,
                            int Z, int Fp, void *stream);
int ddsp_b200_favor_output(const float *qf, const float *ctxT, const float *ctxT_lo, float *out, int B, int H,
                           int F, void *stream);

/* Tensor-pipe microbenchmark of the same kernel: `virtual_tiles` output tiles of 128 x block_n with reduction
 * length K that all read tile 0 of A (128,K) and W (N,K) (operands stay in L2) and store nothing -- the
 * rate at which the TMA -> split -> 3 x tcgen05.mma -> TMEM-drain pipeline runs without HBM traffic.  Used to
 * measure a DFT-32 pass (N = K = 64) as a go / no-go for a tensor-core FFT in the synthesizer kernel. */
int ddsp_b200_tc_microbench(const float *A, const float *W, float *C, int N, int K, int block_n,
                            int virtual_tiles, void *stream);

/* ---- downstream of the synthesizer: enhancer front-end and GUI splice ----------------------
 *
 * ddsp_b200_mel_spectrogram   nsf_hifigan/nvSTFT.py:65-116 (STFT.get_mel, keyshift = 0, speed = 1, center = False):
 *     reflect padding by (win-hop)/2 | (win-hop+1)/2, STFT with the periodic Hann window, sqrt(re^2 + im^2 + 1e-9),
 *     mel_basis (n_mels, n_fft/2+1) product, log(max(., clip_val)).  audio (B,T) -> out (B, n_mels, n_frames),
 *     n_frames = 1 + (T + pads - n_fft) / hop.  n_fft = win_size = 2048 (the 44.1 kHz NSF-HiFiGAN setting).
 *     band_start / band_end (n_mels, int32): first and one-past-last non-zero bin of every mel filter.
 * ddsp_b200_sinc_resample     torchaudio.transforms.Resample(orig, new, lowpass_filter_width) as called at
 *     enhancer.py:47,69 and gui.py:398-401: y[i*nw + j] = sum_k kernel[j][k] * xpad[i*orig + k]; kernel_t is the
 *     table of torchaudio's _get_sinc_resample_kernel TRANSPOSED to (K = 2*width + orig, nw); orig / nw already divided
 *     by their gcd; y (B, T_out), T_out = ceil(nw * T / orig).
 * ddsp_b200_interp_frames     enhancer.py:57-63: out[b,i] = np.interp(dt_out * i, (hop_over_sr * k) / real_factor,
 *     f0[b,k] * scale) in double, ends held; f0 (B,n) view with element strides (fB, fN); out (B, n_out).
 * ddsp_b200_sola_splice       gui.py:408-426 (no phase vocoder): shift = first argmax over d = 0..search of the
 *     normalised cross-correlation of x[d : d+crossfade] with sola_buffer; out[i] = x[shift+i] (crossfaded with the
 *     buffer over the first `crossfade` samples), i < block; sola_buffer <- x[shift+block : shift+block+crossfade];
 *     *shift_out = shift (device memory).  x has n >= block + crossfade + search samples. */
int ddsp_b200_mel_spectrogram(const float *audio, int B, int T, int n_fft, int win_size, int hop,
                              const float *mel_basis, const int *band_start, const int *band_end, int n_mels,
                              float clip_val, float *out, int n_frames, void *stream);
int ddsp_b200_sinc_resample(const float *x, int B, int T, const float *kernel_t, int orig, int nw, int width,
                            float *y, int T_out, void *stream);
int ddsp_b200_interp_frames(const float *f0, int64_t fB, int64_t fN, int B, int n, float scale, double hop_over_sr,
                            double real_factor, double dt_out, float *out, int n_out, void *stream);
int ddsp_b200_sola_splice(const float *x, int n, float *sola_buffer, const float *fade_in, const float *fade_out,
                          int block, int crossfade, int search, float *out, int *shift_out, void *stream);

/* Number of kernel launches the last call of each entry point enqueued on this thread
 * (bench.py reports it as gpu_launches). */
int ddsp_b200_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DDSP_B200_H_ */
