/*
 * afd_b200.h -- C ABI of the B200-native audio-deepfake feature front-end (libafd_b200.so).
 *
 * Drop-in boundary for the feature front-end of gan-police/audiodeepfake-detection ("the reference";
 * all file:line citations are relative to the reference checkout).  The reference has no FFI layer: its
 * boundary is a torch.nn.Module contract (src/audiofakedetect/wavelet_math.py:266-384).  Each entry point
 * below replaces one third-party call the reference makes on that path; the Python modules in
 * audiodeepfake-detection_b200/wavelet_math.py keep the reference's module API and call these through
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device entry points are
 *     asynchronous and stream-ordered; they never synchronise the device.
 *   - the library is stateless (filter taps travel as kernel parameters), hence thread-safe per stream.
 *   - return value: AFD_OK (0), a negative AFD_ERR_* validation code, or a positive cudaError_t.
 *     afd_last_error() returns a thread-local human-readable message for the last non-zero return.
 *   - all arithmetic is IEEE fp32 (FMA), like the reference's fp32 tensors.
 */
#ifndef AFD_B200_H
#define AFD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFD_OK 0
#define AFD_ERR_INVALID_ARG (-1)   /* null pointer, non-positive size, odd / unsupported filter length ... */
#define AFD_ERR_UNSUPPORTED (-2)   /* valid request this build cannot serve (e.g. tree does not fit in shared memory) */
#define AFD_ERR_REFLECT_PAD (-3)   /* a node is not longer than the reflect padding (torch F.pad would raise too) */
#define AFD_ERR_NO_DEVICE (-4)

#define AFD_ORDER_FREQ 0           /* Gray-code / frequency order: ptwt get_level(), pywt order="freq" */
#define AFD_ORDER_NATURAL 1

/* Library version (major*10000 + minor*100 + patch). */
int afd_version(void);

/* Thread-local message for the most recent failing call on this thread ("" if none). */
const char* afd_last_error(void);
/* Build provenance: sha256 (hex) of csrc/ + include/afd_b200.h the library was compiled from -- equals
 * audiodeepfake-detection_b200/build.py:source_hash() when the binary matches the sources on disk; "unknown" for a
 * build that bypassed build.py. */
const char* afd_source_hash(void);

/*
 * Number of coefficients per node after `level` analysis steps on a length-N signal with an F-tap filter:
 * n <- floor((n + F - 1) / 2), `level` times  (ptwt _get_pad / pywt dwt_coeff_len).
 * Replaces: shape inference the reference does by running the transform once (utils.py:589-621).
 */
int afd_wpt_out_len(int64_t N, int F, int level, int64_t* T_out);

/*
 * Fused wavelet-packet feature transform.
 * Replaces: ptwt.WaveletPacket(data, wavelet, mode="reflect") + get_level + per-node loop + torch.stack +
 *           log/abs/pow/sign epilogue  (wavelet_math.py:182-218), i.e. compute_pytorch_packet_representation.
 *
 *   x              device, fp32, B rows of N samples, consecutive rows x_row_stride elements apart
 *   dec_lo_host    HOST pointer to the F low-pass decomposition taps in double precision, as
 *                  pywt.Wavelet(name).dec_lo holds them (wavelet_math.py:239); F even, 2..64.
 *                  The high-pass taps are derived as dec_hi[k] = (-1)^(k+1) dec_lo[F-1-k].
 *   level          tree depth (max_lev), 1..12;  P = 2^level packets
 *   order          AFD_ORDER_FREQ or AFD_ORDER_NATURAL (column order of the packets)
 *   log_scale      0: raw coefficients;  1: log(|c|^power + log_offset)   (reference uses 1e-12)
 *   sign_channel   (loss_less) with log_scale: adds channel 1 = +1 for c >= 0, -1 for c < 0
 *   out            device, fp32, contiguous [B][C][T][P], C = 1 + (log_scale && sign_channel),
 *                  T = afd_wpt_out_len(N, F, level).  The reference hands callers the (0,1,3,2)-permuted VIEW
 *                  of exactly this memory (wavelet_math.py:263).
 *   T_out          optional, receives T.
 */
int afd_wpt_forward(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                    const double* dec_lo_host, int F, int level, int order,
                    float power, int log_scale, float log_offset, int sign_channel,
                    float* out, int64_t* T_out, void* stream);

/*
 * Same transform with the rest of the reference's per-batch bookkeeping fused into the kernel's epilogue.
 * Replaces, in addition to afd_wpt_forward: the per-node WelfordEstimator updates and the block norm of
 * wavelet_math.py:194-203 (data_loader.py:41-63), torchvision Normalize (wavelet_math.py:380-382) and the
 * feature statistics pass of calc_normalization (wavelet_math.py:436-441).  Every pointer below is optional.
 *
 *   node_scale          device fp32 [P]: the coefficients of output column p (frequency / natural order as
 *                       requested) are multiplied by node_scale[p] before the epilogue.  Block norm
 *                       (node / max|node|, :202-203) passes 1 / max|c_p| from a previous node_stats launch.
 *   norm_mean_std_host  HOST fp32 [C][2] = (mean, std) per channel: features leave as (v - mean) / std.
 *   node_stats          device DOUBLE [3][P], accumulated: [0][p] += sum c, [1][p] += sum c^2,
 *                       [2][p] = max([2][p], max|c|) over the B*T RAW coefficients of column p (before
 *                       node_scale).  With the count B*T these give every WelfordEstimator's mean / M2.
 *   feat_moments        device DOUBLE [C][2], accumulated: sum and sum of squares of the features of each channel
 *                       BEFORE normalisation (what calc_normalization feeds its estimator).
 *   out                 may be NULL when node_stats or feat_moments is given: statistics only, no feature
 *                       tensor is written (the first pass of block norm, calc_normalization).
 */
int afd_wpt_forward_ex(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                       const double* dec_lo_host, int F, int level, int order,
                       float power, int log_scale, float log_offset, int sign_channel,
                       const float* node_scale, const float* norm_mean_std_host,
                       double* node_stats, double* feat_moments,
                       float* out, int64_t* T_out, void* stream);

/*
 * Same transform with HOST input and output buffers (pinned memory recommended): frames are streamed through
 * the device in chunks with H2D copy, kernel and D2H copy overlapped on separate streams.  Blocks until `out_host`
 * is complete.  This is the call a CPU-side user of the reference (numpy in, numpy out) would make.
 */
int afd_wpt_forward_host(const float* x_host, int64_t B, int64_t N, int64_t x_row_stride,
                         const double* dec_lo_host, int F, int level, int order,
                         float power, int log_scale, float log_offset, int sign_channel,
                         float* out_host, int64_t* T_out, int device, int64_t chunk_frames);

/*
 * Fused STFT power spectrogram.
 * Replaces: torchaudio.transforms.Spectrogram(n_fft, hop_length, power)(x) -> torch.stft(center=True,
 *           pad_mode="reflect", window=hann(periodic), onesided) -> abs().pow(power), and the optional
 *           log(spec + 1e-12)  (wavelet_math.py:47,63-66).
 *   out   device, fp32, contiguous [B][1][frames][n_fft/2+1], frames = 1 + (N + 2*(n_fft/2) - n_fft) / hop.  The caller exposes the
 *         (0,1,3,2)-permuted view [B,1,bins,frames] the reference returns.
 */
int afd_stft_power(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                   int n_fft, int hop, float power, int log_scale, float log_offset,
                   float* out, void* stream);

/*
 * Same transform with torchvision Normalize (wavelet_math.py:380-382) and the feature statistics of
 * calc_normalization (wavelet_math.py:436-441) fused into the epilogue; both pointers optional.
 *   norm_mean_std_host  HOST fp32 [2] = (mean, std): features leave as (v - mean) / std.
 *   feat_moments        device DOUBLE [2], accumulated: sum and sum of squares of the features BEFORE normalisation.
 *   out                 may be NULL when feat_moments is given (statistics only).
 */
int afd_stft_power_ex(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                      int n_fft, int hop, float power, int log_scale, float log_offset,
                      const float* norm_mean_std_host, double* feat_moments, float* out, void* stream);

/* frames = 1 + (N + 2*(n_fft/2) - n_fft) / hop  (torch.stft, center=True), bins = n_fft / 2 + 1 */
int afd_stft_out_shape(int64_t N, int n_fft, int hop, int64_t* frames, int64_t* bins);

int afd_stft_power_host(const float* x_host, int64_t B, int64_t N, int64_t x_row_stride,
                        int n_fft, int hop, float power, int log_scale, float log_offset,
                        float* out_host, int device, int64_t chunk_frames);

/*
 * Haar wavelet-packet "fingerprint" accumulation.
 * Replaces: pywt.WaveletPacket(clips, "haar", mode="reflect").get_level(level, order="freq") + np.stack +
 *           np.abs + the sum part of np.mean  (scripts/freq_visual/fingerprints.py:101-115).
 *   sums   device, DOUBLE [2^level]; sums[p] += sum over the B clips and all positions of |c_p|.
 *   count  optional device int64 scalar; += B * T  (the number of terms behind each sum), so that
 *          mean = sums / count also after an all-reduce(sum) of both across ranks.
 */
int afd_haar_fingerprint_accum(const float* x, int64_t B, int64_t N, int64_t x_row_stride, int level,
                               double* sums, int64_t* count, void* stream);

int afd_haar_fingerprint_host(const float* x_host, int64_t B, int64_t N, int64_t x_row_stride, int level,
                              double* sums_host, int64_t* count_host, int device, int64_t chunk_frames);

/*
 * Mean-spectrum ("rFFT") fingerprint, part 1: per-sample sums over clips.
 * Replaces: np.fft.rfft(clip_array) -> np.fft.irfft -> the sum part of np.mean(..., 0)
 *           (scripts/freq_visual/fingerprints.py:51-59).  rfft -> irfft over all bins is the identity and the mean
 *           is linear, so only the column sums of the clips are needed: one coalesced pass, nothing written.
 *   sums   device DOUBLE [N]; sums[n] += sum over the B clips of x[b][n]
 *   count  optional device int64 scalar; += B   (mean clip = sums / count, also after an all-reduce of both)
 */
int afd_clip_sum_accum(const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                       double* sums, int64_t* count, void* stream);

/*
 * Mean-spectrum fingerprint, part 2: mag[k] = | sum_n scale * x[n] * exp(-2 pi i k n / N) |, k = 0 .. N/2.
 * Replaces: np.abs(np.fft.rfft(masked_time_mean))  (fingerprints.py:60); x = the clip sums, scale = 1 / count.
 * One direct real DFT in double precision with exact phase reduction (any N up to 2^24; N = 22050 has no
 * power-of-two path and the transform runs once per fingerprint).  x: device DOUBLE [N]; mag: device DOUBLE [N/2+1].
 */
int afd_rdft_magnitude(const double* x, int64_t N, double scale, double* mag, void* stream);

/*
 * Sample-rate conversion in front of the frame cutter.
 * Replaces: torchaudio.functional.resample(audio, sample_rate, self.resample_rate)  (src/audiofakedetect/data_loader.py:341-344;
 *           torchaudio defaults: sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) -- the polyphase FIR of torchaudio's
 *           _get_sinc_resample_kernel / _apply_sinc_resample_kernel with the tap table built once per (orig, new) pair.
 *   x    device fp32 [B][n_in] (row stride x_row_stride);  out  device fp32 [B][n_out] (row stride out_row_stride),
 *   n_out = ceil(new * n_in / orig) after reducing the rates by their gcd (afd_resample_out_len).  B <= 65535 rows per call.
 */
int afd_resample_out_len(int64_t n_in, int orig_freq, int new_freq, int64_t* n_out);
int afd_resample(const float* x, int64_t B, int64_t n_in, int64_t x_row_stride, int orig_freq, int new_freq,
                 float* out, int64_t out_row_stride, void* stream);

/*
 * Diagnostic: the paraunitary lattice (plane rotations + unit delays) the wavelet-packet kernel evaluates instead
 * of the direct-form filter pair when `usable` is 1 -- half the multiplies per coefficient pair.  tan_theta
 * receives F/2 stage tangents (stage 0 first), scale the product of the stage cosines, residual the largest
 * |tap error| of the re-synthesised filter pair against dec_lo / dec_hi.  Host only; no device work.
 */
int afd_wpt_lattice_info(const double* dec_lo_host, int F, double* tan_theta, double* scale,
                         double* residual, int* usable);

/*
 * Diagnostic: the launch configuration afd_wpt_forward would use for (N, taps, level): dynamic shared memory per
 * CTA, resident CTAs per SM, whether the lattice path is taken, and per barrier-delimited pass after level 1 the
 * number of work items and the item size.  pass_items / pass_r must hold 24 ints.  Host only; no device work.
 */
int afd_wpt_plan_info(int64_t N, const double* dec_lo_host, int F, int level, int* smem_bytes, int* ctas_per_sm,
                      int* lattice, int* passes, int* pass_items, int* pass_r);

/*
 * Measurement helper (bench.py): runs `iters` back-to-back launches of an FFMA-only kernel that fills every
 * SM and returns the achieved fp32 TFLOP/s -- the live FP32-FMA roofline denominator.
 */
int afd_measure_fp32_fma_tflops(int iters, double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFD_B200_H */
