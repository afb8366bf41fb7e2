// Entry points not built yet in this round return B200VQA_ENOTLOADED (removed as they land).
#include "context.h"
namespace b200vqa {
void free_resnet(ResNetWeights*) {}
void free_vit(ViTWeights*) {}
void free_head(HeadWeights*) {}
}
extern "C" {
int b200vqa_load_resnet50(b200vqa_t*, int, const char* const*, const float* const*, const int64_t*) { return B200VQA_ENOTLOADED; }
int b200vqa_load_vitb16(b200vqa_t*, int, const char* const*, const float* const*, const int64_t*) { return B200VQA_ENOTLOADED; }
int b200vqa_load_head(b200vqa_t*, int, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const double*, const double*, const double*) { return B200VQA_ENOTLOADED; }
int b200vqa_resnet50_features(b200vqa_t*, const uint8_t*, int, int, float*, float*, void*) { return B200VQA_ENOTLOADED; }
int b200vqa_vitb16_features(b200vqa_t*, const uint8_t*, int, int, float*, void*) { return B200VQA_ENOTLOADED; }
int b200vqa_temporal_mean_concat(const float*, const float*, const float*, const float*, const float*, const float*, const int32_t*, const int32_t*, int, float*, void*) { return B200VQA_ENOTLOADED; }
int b200vqa_head_forward(b200vqa_t*, const float*, int, float*, void*) { return B200VQA_ENOTLOADED; }
}
