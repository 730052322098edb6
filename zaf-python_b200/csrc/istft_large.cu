// ISTFT warp kernels for window lengths 2048 (BASELINE cfg 2, with the masked variants) and 4096; see istft.cu.
#include "istft_kernels.cuh"

namespace zafb {

int istft_warp_dispatch_large(const zafb_stft_plan* p, const float2* s2, int64_t clips, int64_t nt, float* yy, int64_t y_stride,
                              cudaStream_t st, int64_t pitch, int onesided, const float* mask, int64_t mask_pitch) {
    const int n = int(p->n);
    const int64_t ratio = (p->hop > 0 && n % p->hop == 0) ? n / p->hop : 0;
    if (n == 4096) {
        if (ratio == 2) return launch_istft_warp<4096, 2>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
        if (ratio == 4) return launch_istft_warp<4096, 4>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
        return launch_istft_warp<4096, 8>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
    }
    if (ratio == 2) return launch_istft_warp<2048, 2>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
    if (ratio == 4) return launch_istft_warp<2048, 4>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
    return launch_istft_warp<2048, 8>(p, s2, clips, nt, yy, y_stride, st, pitch, onesided, mask, mask_pitch);
}

int istft_binmajor_dispatch_large(const zafb_stft_plan* p, const float2* s2, int64_t n_clips, int64_t nt, float* y, int64_t y_stride,
                                  cudaStream_t st) {
    return p->n / p->hop == 2 ? launch_istft_binmajor<2048, 2>(p, s2, n_clips, nt, y, y_stride, st)
                              : launch_istft_binmajor<2048, 4>(p, s2, n_clips, nt, y, y_stride, st);
}

}  // namespace zafb
