// Entry points declared in include/zafb200.h whose kernels are not built yet.
#include "common.cuh"
using namespace zafb;
#define NOT_YET(name) return fail(ZAFB_E_UNSUPPORTED, name ": not implemented yet")
extern "C" {
int zafb_mdct_plan_create(zafb_mdct_plan**, const double*, int64_t) { NOT_YET("zafb_mdct_plan_create"); }
int zafb_mdct_plan_destroy(zafb_mdct_plan*) { return ZAFB_OK; }
int zafb_mdct_f32(const zafb_mdct_plan*, const float*, int64_t, int64_t, int64_t, float*, int, void*) { NOT_YET("zafb_mdct_f32"); }
int zafb_imdct_f32(const zafb_mdct_plan*, const float*, int64_t, int64_t, int, float*, int64_t, void*) { NOT_YET("zafb_imdct_f32"); }
int zafb_dct_plan_create(zafb_dct_plan**, int, int, int64_t) { NOT_YET("zafb_dct_plan_create"); }
int zafb_dct_plan_destroy(zafb_dct_plan*) { return ZAFB_OK; }
int zafb_dct_f32(const zafb_dct_plan*, const float*, int64_t, int64_t, float*, int64_t, void*) { NOT_YET("zafb_dct_f32"); }
int zafb_mel_plan_create(zafb_mel_plan**, const double*, int64_t, int64_t, const double*, int64_t, int64_t) { NOT_YET("zafb_mel_plan_create"); }
int zafb_mel_plan_destroy(zafb_mel_plan*) { return ZAFB_OK; }
int zafb_melspectrogram_f32(const zafb_mel_plan*, const float*, int64_t, int64_t, int64_t, float*, int, void*) { NOT_YET("zafb_melspectrogram_f32"); }
int zafb_mfcc_f32(const zafb_mel_plan*, const float*, int64_t, int64_t, int64_t, float*, int, void*) { NOT_YET("zafb_mfcc_f32"); }
int zafb_cqt_plan_create(zafb_cqt_plan**, int64_t, int64_t, const int32_t*, const int32_t*, const double*, int64_t) { NOT_YET("zafb_cqt_plan_create"); }
int zafb_cqt_plan_destroy(zafb_cqt_plan*) { return ZAFB_OK; }
int zafb_cqt_f32(const zafb_cqt_plan*, const float*, int64_t, int64_t, int64_t, int64_t, float*, int, void*) { NOT_YET("zafb_cqt_f32"); }
}
