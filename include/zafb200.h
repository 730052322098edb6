/*
 * zafb200.h -- C ABI of libzafb200.so: the B200 (sm_100a) implementation of the
 * data-parallel transform hot path of zafarrafii/Zaf-Python (zaf.py).
 *
 * The reference has no FFI of its own: its boundary is the module-level Python
 * signatures.  Each compute entry point below names the
 * reference function (file:line in /root/reference) whose body it replaces; the
 * ctypes binding that calls it lives in zaf-python_b200/_lib.py and the drop-in
 * functions with the reference's signatures in zaf-python_b200/__init__.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no exceptions cross the boundary.
 *   - every function returns 0 (ZAFB_OK) or a negative ZAFB_E_* code;
 *     zafb_last_error() returns a thread-local message for the last failure.
 *   - the caller owns every data buffer; the library owns only opaque plans
 *     (immutable after creation: window / twiddle / filterbank tables in HBM).
 *   - "_f32" entry points take DEVICE pointers and a CUDA stream (cudaStream_t
 *     passed as void*; NULL = default stream) and never synchronise.
 *   - "_host_f32" entry points take HOST pointers, do H2D -> kernels -> D2H in
 *     pipelined chunks on the library's own streams and return when the output
 *     is complete.
 *   - complex data is interleaved (re, im) float32.
 *   - arithmetic type: fp32 (tables are generated in float64 and rounded once).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails
 *     with ZAFB_E_CUDA.
 */
#ifndef ZAFB200_H
#define ZAFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZAFB_OK 0
#define ZAFB_E_BADARG (-1)      /* -> ValueError            */
#define ZAFB_E_UNSUPPORTED (-2) /* -> NotImplementedError   */
#define ZAFB_E_CUDA (-3)        /* -> RuntimeError          */
#define ZAFB_E_NOMEM (-4)       /* -> MemoryError           */
#define ZAFB_E_NCCL (-5)        /* -> RuntimeError          */

/* Output layouts of the (frequency, time) matrices.
 *   ZAFB_LAYOUT_FRAME_MAJOR: memory is [clip][frame][bin]   (bin contiguous; the Python
 *                            wrapper exposes it as a transposed view of shape (bins, frames))
 *   ZAFB_LAYOUT_BIN_MAJOR  : memory is [clip][bin][frame]   (frame contiguous; the reference's
 *                            C-order (window_length, number_times) array, zaf.py:128) */
#define ZAFB_LAYOUT_FRAME_MAJOR 0
#define ZAFB_LAYOUT_BIN_MAJOR 1

/* ---------------------------------------------------------------- runtime */
const char* zafb_last_error(void);
const char* zafb_version(void);
int zafb_device_count(int* count);
int zafb_init(int device);              /* cudaSetDevice + context warm-up            */
int zafb_shutdown(void);                /* release the host-pipeline streams/buffers   */
int zafb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                     size_t* total_mem, char* name, size_t name_len);

int zafb_malloc(void** dev_ptr, size_t bytes);
int zafb_free(void* dev_ptr);
int zafb_host_alloc(void** host_ptr, size_t bytes);   /* pinned */
int zafb_host_free(void* host_ptr);
int zafb_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int zafb_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int zafb_memcpy_d2d(void* dst_dev, const void* src_dev, size_t bytes, void* stream);
/* rows of width_bytes at different pitches; kind: 0 = host to device, 1 = device to host, 2 = device to device */
int zafb_memcpy2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                  size_t rows, int kind, void* stream);
int zafb_memset(void* dev_ptr, int value, size_t bytes, void* stream);

int zafb_stream_create(void** stream);
int zafb_stream_destroy(void* stream);
int zafb_stream_sync(void* stream);
int zafb_device_sync(void);
int zafb_event_create(void** event);
int zafb_event_destroy(void* event);
int zafb_event_record(void* event, void* stream);
int zafb_event_sync(void* event);
/* the work queued on `stream` after this call starts only when `event` (recorded on another stream) has completed */
int zafb_stream_wait_event(void* stream, void* event);
int zafb_event_elapsed_ms(void* start, void* stop, float* ms);
/* zaf.wavread's normalisation (zaf.py:1199-1202) on the device: interleaved (frame, channel) int16 PCM ->
 * planar fp32 [channel][frame] (row pitch out_stride), x / 2^15 exactly; mono != 0 writes the channel mean instead. */
int zafb_pcm16_to_f32(const int16_t* pcm_dev, int64_t frames, int channels, int mono, float* out_dev,
                      int64_t out_stride, void* stream);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t zafb_launch_count(void);
/* Bytes the *_host_f32 pipelines have copied host->device and device->host since the library was loaded. */
int zafb_host_copy_bytes(int64_t* h2d, int64_t* d2h);

/* ------------------------------------------------------- integer bookkeeping
 * Bit-exact restatements of the reference's frame arithmetic (no device needed). */
/* zaf.py:99-121   pad = floor(N/2); nt = ceil((ns+2*pad-N)/hop)+1; tail pad          */
int zafb_stft_geometry(int64_t number_samples, int64_t window_length, int64_t step_length,
                       int64_t* pad, int64_t* number_times, int64_t* tail);
/* zaf.py:217,236-238  OLA length nt*hop+(N-hop); output = ola[N-hop : ola-(N-hop)] with Python
 * slice semantics; `trim` receives the slice start (N-hop whenever hop <= N)            */
int zafb_istft_geometry(int64_t window_length, int64_t number_times, int64_t step_length,
                        int64_t* ola_length, int64_t* trim, int64_t* number_samples);
/* zaf.py:1029-1041 */
int zafb_mdct_geometry(int64_t number_samples, int64_t window_length, int64_t* half,
                       int64_t* number_times, int64_t* tail);
/* zaf.py:1132,1182   OLA length M*(nt+1); output = [M : -M-1]                        */
int zafb_imdct_geometry(int64_t number_frequencies, int64_t number_times, int64_t* ola_length,
                        int64_t* number_samples);
/* zaf.py:603-620; `step` is computed by the caller (Python round-half-even of a float) */
int zafb_cqt_geometry(int64_t number_samples, int64_t step_length, int64_t fft_length,
                      int64_t* number_times, int64_t* front_pad, int64_t* back_pad);

/* ------------------------------------------------------------- STFT / ISTFT
 * Plan: window (float64 on host, rounded once to fp32), twiddle tables. */
typedef struct zafb_stft_plan zafb_stft_plan;
int zafb_stft_plan_create(zafb_stft_plan** plan, const double* window, int64_t window_length,
                          int64_t step_length);
int zafb_stft_plan_destroy(zafb_stft_plan* plan);

/* Replaces the body of zaf.stft (zaf.py:95-141) for a batch of n_clips signals of ns samples
 * (clip c starts at x + c*clip_stride).  out: n_clips * window_length * number_times complex64
 * in `layout`; the full two-sided spectrum, like the reference. */
int zafb_stft_f32(const zafb_stft_plan* plan, const float* x, int64_t n_clips, int64_t ns,
                  int64_t clip_stride, float* out, int layout, void* stream);
/* Replaces zaf.istft (zaf.py:214-243).  spec: n_clips * window_length * nt complex64 in `layout`;
 * y: n_clips rows of nt*hop-(N-hop) samples, row c at y + c*y_stride. */
int zafb_istft_f32(const zafb_stft_plan* plan, const float* spec, int64_t n_clips, int64_t nt,
                   int layout, float* y, int64_t y_stride, void* stream);
/* Host-buffer versions (pinned or pageable host memory; chunked and pipelined).
 * zafb_stft_host_f32, results of 256 MB and more (either layout): the spectrum of a real signal is Hermitian, so only bins
 * 0 .. N/2 of each frame cross PCIe and host threads write the mirrored half, X[N-k] = conj(X[k]) -- exact, the same bits
 * the device kernel stores (ZAFB_HOST_MIRROR=0 copies the full spectrum instead; ZAFB_HOST_MIRROR_THREADS sets the
 * thread count; default min(16, host cores / LOCAL_WORLD_SIZE), and the path stays off below 6 threads). */
int zafb_stft_host_f32(const zafb_stft_plan* plan, const float* x_host, int64_t n_clips, int64_t ns,
                       int64_t clip_stride, float* out_host, int layout);
int zafb_istft_host_f32(const zafb_stft_plan* plan, const float* spec_host, int64_t n_clips,
                        int64_t nt, int layout, float* y_host, int64_t y_stride);

/* One-sided spectra -- an explicit NON-reference extension (zaf.stft always returns the two-sided
 * spectrum, zaf.py:139): bins 0 .. floor(N/2) of every frame, FRAME_MAJOR, `pitch` complex64 elements between consecutive
 * frames (pitch >= N/2 + 1; pitch == N lets a peer write the lower half of a two-sided buffer in place).  The warp kernels
 * store / load the half directly -- half the HBM traffic of the spectrum; zafb_spec_mirror_f32 rebuilds the two-sided form
 * (dst[k] = src[k], dst[N-k] = conj(src[k]); src == dst with src_pitch == N fills the upper half in place). */
int zafb_stft_onesided_f32(const zafb_stft_plan* plan, const float* x, int64_t n_clips, int64_t ns,
                           int64_t clip_stride, float* out, int64_t out_pitch, void* stream);
int zafb_istft_onesided_f32(const zafb_stft_plan* plan, const float* spec, int64_t n_clips, int64_t nt,
                            int64_t spec_pitch, float* y, int64_t y_stride, void* stream);
int zafb_spec_mirror_f32(const float* src, int64_t src_pitch, int64_t frames, int64_t window_length,
                         float* dst, void* stream);
/* zaf.istft of np.concatenate((mask, mask[-2:0:-1])) * X (zaf.py:185-190) with the multiply fused into the ISTFT's loads:
 * `mask` holds a real value for bins 0 .. N/2 of every frame (mask_pitch floats per frame); spec is FRAME_MAJOR, two-sided
 * (onesided = 0, spec_pitch = N) or one-sided (bins 0 .. N/2, spec_pitch complex elements per frame).  The fused kernel
 * exists for window_length 2048, hop 512; other geometries return ZAFB_E_UNSUPPORTED (multiply with zafb_spec_mask_f32
 * first). */
int zafb_istft_masked_f32(const zafb_stft_plan* plan, const float* spec, int64_t n_clips, int64_t nt, int64_t spec_pitch,
                          int onesided, const float* mask, int64_t mask_pitch, float* y, int64_t y_stride, void* stream);

/* The host-side half of that path, usable on its own: given `frames` frame-major frames of `window_length` complex64
 * bins (window_length a multiple of 4) whose bins 0 .. N/2 are valid, writes bins N/2+1 .. N-1 as conj of bins
 * N/2-1 .. 1 -- the two-sided spectrum zaf.stft returns (zaf.py:139) from a one-sided one.  No device involved. */
int zafb_host_mirror_fill(float* spectrum_host, int64_t frames, int64_t window_length);

/* ------------------------------------------------------------- MDCT / IMDCT */
typedef struct zafb_mdct_plan zafb_mdct_plan;
int zafb_mdct_plan_create(zafb_mdct_plan** plan, const double* window, int64_t window_length);
int zafb_mdct_plan_destroy(zafb_mdct_plan* plan);
/* zaf.mdct (zaf.py:1025-1075): out is n_clips * (N/2) * nt float32 in `layout`.  ZAFB_LAYOUT_BIN_MAJOR is the (N/2, nt)
 * C-order array the reference returns (zaf.py:1073); for N = 2048 / 1024 one kernel writes it directly, at any 4-byte
 * phase of `out` (x rows 8-byte aligned); other lengths go through frame-major scratch and a tiled transpose. */
int zafb_mdct_f32(const zafb_mdct_plan* plan, const float* x, int64_t n_clips, int64_t ns,
                  int64_t clip_stride, float* out, int layout, void* stream);
/* zaf.imdct (zaf.py:1125-1184): y rows of M*(nt-1)-1 samples.  BIN_MAJOR input (the array zaf.mdct returns, zaf.py:1159)
 * is read directly for N = 2048 / 1024 when n_clips is at least half the SM count (one CTA walks one clip). */
int zafb_imdct_f32(const zafb_mdct_plan* plan, const float* spec, int64_t n_clips, int64_t nt,
                   int layout, float* y, int64_t y_stride, void* stream);

/* Host-buffer versions of the two calls above (same chunked pipeline as zafb_stft_host_f32). */
int zafb_mdct_host_f32(const zafb_mdct_plan* plan, const float* x_host, int64_t n_clips, int64_t ns,
                       int64_t clip_stride, float* out_host, int layout);
int zafb_imdct_host_f32(const zafb_mdct_plan* plan, const float* spec_host, int64_t n_clips,
                        int64_t nt, int layout, float* y_host, int64_t y_stride);

/* ---------------------------------------------------------------- DCT / DST
 * zaf.dct (zaf.py:759-839) / zaf.dst (zaf.py:901-981): orthonormal types 1..4 of `batch`
 * vectors of length n (vector b at x + b*stride).  kind: 0 = DCT, 1 = DST. */
typedef struct zafb_dct_plan zafb_dct_plan;
int zafb_dct_plan_create(zafb_dct_plan** plan, int kind, int type, int64_t n);
int zafb_dct_plan_destroy(zafb_dct_plan* plan);
int zafb_dct_f32(const zafb_dct_plan* plan, const float* x, int64_t batch, int64_t stride,
                 float* out, int64_t out_stride, void* stream);

int zafb_dct_host_f32(const zafb_dct_plan* plan, const float* x_host, int64_t batch, int64_t stride,
                      float* out_host, int64_t out_stride);

/* ----------------------------------------------------- mel spectrogram / MFCC
 * The filterbank is passed dense row-major (n_mels x N/2, float64) -- the wrapper calls
 * .toarray() exactly like zaf.py:373 -- and packed into per-row bands at plan creation.
 * n_coef = 0 builds a melspectrogram-only plan. */
typedef struct zafb_mel_plan zafb_mel_plan;
int zafb_mel_plan_create(zafb_mel_plan** plan, const double* window, int64_t window_length,
                         int64_t step_length, const double* filterbank, int64_t n_mels,
                         int64_t n_coef);
int zafb_mel_plan_destroy(zafb_mel_plan* plan);
/* How the filterbank is applied.  FUSED (default): banded multiply-add inside the STFT kernel, the
 * spectrum never reaches HBM.  TENSOR: the dense contraction of zaf.py:373/445 (np.matmul with the
 * densified filterbank) on the tcgen05 tensor cores in 3xTF32 (window_length 1024 only). */
#define ZAFB_MEL_ROUTE_FUSED 0
#define ZAFB_MEL_ROUTE_TENSOR 1
int zafb_mel_plan_set_route(zafb_mel_plan* plan, int route);
/* Arithmetic of the spectrum, filterbank sums and logarithm: 32 (default) or 64 bits.  The float64 route (any
 * power-of-two window length) exists for purely tonal material, where the LOG in mfcc amplifies the fp32 floor of the
 * FFT; pass the float64 window again.  Inputs and outputs stay fp32. */
int zafb_mel_plan_set_precision(zafb_mel_plan* plan, int bits, const double* window);
/* zaf.melspectrogram (zaf.py:369-375): out n_clips * n_mels * nt float32 in `layout`. */
int zafb_melspectrogram_f32(const zafb_mel_plan* plan, const float* x, int64_t n_clips, int64_t ns,
                            int64_t clip_stride, float* out, int layout, void* stream);
/* zaf.mfcc (zaf.py:436-454): out n_clips * n_coef * nt float32 in `layout`. */
int zafb_mfcc_f32(const zafb_mel_plan* plan, const float* x, int64_t n_clips, int64_t ns,
                  int64_t clip_stride, float* out, int layout, void* stream);

int zafb_melspectrogram_host_f32(const zafb_mel_plan* plan, const float* x_host, int64_t n_clips,
                                 int64_t ns, int64_t clip_stride, float* out_host, int layout);
int zafb_mfcc_host_f32(const zafb_mel_plan* plan, const float* x_host, int64_t n_clips, int64_t ns,
                       int64_t clip_stride, float* out_host, int layout);

/* ------------------------------------------------------- CQT spectrogram/chroma
 * The kernel is passed as CSR (complex128 data as interleaved doubles), i.e. the
 * scipy.sparse.csr_matrix zaf.cqtkernel returns (zaf.py:554-557). */
typedef struct zafb_cqt_plan zafb_cqt_plan;
int zafb_cqt_plan_create(zafb_cqt_plan** plan, int64_t n_freqs, int64_t fft_length,
                         const int32_t* indptr, const int32_t* indices, const double* data_ri,
                         int64_t step_length);
int zafb_cqt_plan_destroy(zafb_cqt_plan* plan);
/* How the kernel is applied.  FUSED (default): banded multiply-add inside the FFT kernel.  TENSOR: the spectrum's
 * (Re, Im) rows against the dense real kernel block on the tcgen05 tensor cores in 3xTF32 (fft_length 32768, real
 * kernels whose bands stay below fft_length/2 -- every kernel zaf.cqtkernel builds). */
#define ZAFB_CQT_ROUTE_FUSED 0
#define ZAFB_CQT_ROUTE_TENSOR 1
int zafb_cqt_plan_set_route(zafb_cqt_plan* plan, int route);
/* zaf.cqtspectrogram (zaf.py:603-635): out n_clips * n_freqs * nt float32 in `layout`.
 * octave_resolution > 0 additionally folds rows i::octave_resolution (zaf.cqtchromagram,
 * zaf.py:682-700) and the output has octave_resolution rows instead. */
int zafb_cqt_f32(const zafb_cqt_plan* plan, const float* x, int64_t n_clips, int64_t ns,
                 int64_t clip_stride, int64_t octave_resolution, float* out, int layout,
                 void* stream);

int zafb_cqt_host_f32(const zafb_cqt_plan* plan, const float* x_host, int64_t n_clips, int64_t ns,
                      int64_t clip_stride, int64_t octave_resolution, float* out_host, int layout);

/* ------------------------------------------------- device-resident stages between transforms
 * SURVEY.md section 8f-3.  The reference's own demos are chains -- stft -> time-frequency mask -> istft (zaf.py:162-198),
 * mdct -> coefficients processed -> imdct (zaf.py:1098-1105) -- whose middle step is NumPy elementwise arithmetic.  These
 * entry points are that arithmetic on DEVICE pointers, so a chain never crosses PCIe.  `layout` as above; spectra are
 * n_clips * bins * frames complex64, masks / magnitudes float32 in the same layout. */
/* |X[k]| of bins 0 .. keep_bins-1 (zaf.py:176: abs(audio_stft[0:number_frequencies, :])) */
int zafb_spec_abs_f32(const float* spec, int64_t n_clips, int64_t bins, int64_t frames, int layout,
                      int64_t keep_bins, float* out, void* stream);
/* out = spec * mask; mask_bins == bins, or bins/2 + 1 with the mask mirrored onto bins N-k like
 * np.concatenate((mask, mask[-2:0:-1, :])) (zaf.py:185).  out may alias spec. */
int zafb_spec_mask_f32(const float* spec, int64_t n_clips, int64_t bins, int64_t frames, int layout,
                       const float* mask, int64_t mask_bins, float* out, void* stream);
int zafb_ratio_min_f32(const float* a, const float* b, int64_t n, float* out, void* stream); /* min(a,b)/a, zaf.py:181 */
int zafb_mul_f32(const float* a, const float* b, int64_t n, float* out, void* stream);       /* a*b */
int zafb_quantize_f32(const float* x, int64_t n, float step, float* out, void* stream);      /* step*rint(x/step) */
/* number of differing 32-bit words (bitwise check of a sharded result against the unsharded one); synchronises */
int zafb_count_mismatch_u32(const void* a, const void* b, int64_t n_words, int64_t* count, void* stream);

/* ------------------------------------------------- multi-GPU batch split / merge
 * One process per GPU.  The reference has no distributed code; clips are
 * independent, so the only collectives on the path move a batch: rows (a clip, or one clip's
 * spectrogram; opaque byte strings) are split from / merged on a root rank over NCCL (NVLink 5 /
 * NVSwitch).  Rank r of R owns rows [floor(r n / R), floor((r+1) n / R)).  NCCL is loaded with
 * dlopen at the first call; every pointer is a DEVICE pointer, transfers are stream-ordered. */
typedef struct zafb_comm zafb_comm;
int zafb_dist_shard_range(int64_t n_rows, int rank, int world, int64_t* begin, int64_t* end);
int zafb_dist_nccl_version(int* version);
int zafb_dist_unique_id(void* id128);                 /* rank 0: 128-byte NCCL id to hand to its peers */
int zafb_dist_init(zafb_comm** comm, const void* id128, int rank, int world); /* after zafb_init(device) */
int zafb_dist_destroy(zafb_comm* comm);
int zafb_dist_rank(const zafb_comm* comm, int* rank, int* world);
int zafb_dist_broadcast(zafb_comm* comm, void* buf, size_t bytes, int root, void* stream);
int zafb_dist_scatter_rows(zafb_comm* comm, const void* src_root, void* dst, int64_t n_rows,
                           int64_t row_bytes, int root, void* stream);
int zafb_dist_gather_rows(zafb_comm* comm, const void* src, void* dst_root, int64_t n_rows,
                          int64_t row_bytes, int root, void* stream);
int zafb_dist_allgather_rows(zafb_comm* comm, const void* src, void* dst, int64_t n_rows,
                             int64_t row_bytes, void* stream);
int zafb_dist_max_f64(zafb_comm* comm, double* value, void* stream);  /* max over ranks, synchronises */
/* Peer memory (CUDA IPC, same node): the root exports a cudaMalloc'ed buffer (the START of the
 * allocation), peers map it and use the mapped address as the output pointer of any *_f32 entry
 * point: the transform's own stores cross NVLink, no separate gather. */
int zafb_dist_peer_export(const void* dev_ptr, void* handle64);
int zafb_dist_peer_open(const void* handle64, void** mapped);
int zafb_dist_peer_close(void* mapped);

#ifdef __cplusplus
}
#endif
#endif /* ZAFB200_H */
