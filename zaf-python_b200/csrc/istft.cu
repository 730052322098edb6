// ISTFT warp kernels (zaf.py:144-243) for window lengths 256, 512, 1024 and the dispatch over window length / hop ratio.
//   istft_warp_kernel<N, R, WARPS, ONESIDED, MASKED>   frame-major spectra, one warp per run of hop-blocks
//   istft_binmajor_kernel<N, R>                        the reference's C-order memory read directly
// The kernels are templates in istft_kernels.cuh; N = 2048 / 4096 are instantiated in istft_large.cu (the two files
// compile in parallel -- as one translation unit with stft.cu they were a 6-minute ptxas run).  Entry points, plans and
// the generic kernels live in stft.cu.
#include "istft_kernels.cuh"

namespace zafb {

int istft_warp_dispatch_large(const zafb_stft_plan* p, const float2* s2, int64_t clips, int64_t nt, float* yy, int64_t y_stride,
                              cudaStream_t st, int64_t pitch, int onesided, const float* mask, int64_t mask_pitch);  // istft_large.cu
int istft_binmajor_dispatch_large(const zafb_stft_plan* p, const float2* s2, int64_t n_clips, int64_t nt, float* y, int64_t y_stride,
                                  cudaStream_t st);

int istft_warp_dispatch(const zafb_stft_plan* p, const float2* s2, int64_t clips, int64_t nt, float* yy, int64_t y_stride,
                        cudaStream_t st, int64_t pitch, int onesided, const float* mask, int64_t mask_pitch) {
    const int n = int(p->n);
    const int64_t ratio = (p->hop > 0 && n % p->hop == 0) ? n / p->hop : 0;
    if (n >= 2048) return istft_warp_dispatch_large(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
    if (n == 1024) {
        if (ratio == 2) return launch_istft_warp<1024, 2>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
        if (ratio == 4) return launch_istft_warp<1024, 4>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
        return launch_istft_warp<1024, 8>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
    }
    if (n == 256) {
        if (ratio == 2) return launch_istft_warp<256, 2>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
        return launch_istft_warp<256, 4>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
    }
    if (ratio == 2) return launch_istft_warp<512, 2>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
    if (ratio == 4) return launch_istft_warp<512, 4>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
    return launch_istft_warp<512, 8>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
}

int istft_binmajor_dispatch(const zafb_stft_plan* p, const float2* s2, int64_t n_clips, int64_t nt, float* y, int64_t y_stride,
                            cudaStream_t st) {
    if (p->n == 2048) return istft_binmajor_dispatch_large(p, s2, n_clips, nt, y, y_stride, st);
    return p->n / p->hop == 2 ? launch_istft_binmajor<1024, 2>(p, s2, n_clips, nt, y, y_stride, st)
                              : launch_istft_binmajor<1024, 4>(p, s2, n_clips, nt, y, y_stride, st);
}

}  // namespace zafb
