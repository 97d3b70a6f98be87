// Entry points declared in include/bore_b200.h whose kernels are not written yet.
// Each fails loudly; nothing here computes on the CPU.
#include "common.cuh"
extern "C" {
int bore_mlp_fit(bore_mlp *, int, int, const float *, const float *, int, int, int, int,
                 const int32_t *, int, float, float *, void *) {
  bore_set_error("bore_mlp_fit: not implemented yet"); return -1; }
int bore_mlp_evaluate(bore_mlp *, int, const float *, const float *, int, float, float *, void *) {
  bore_set_error("bore_mlp_evaluate: not implemented yet"); return -1; }
size_t bore_lbfgsb_workspace_bytes(int, int, int) { return 0; }
int bore_lbfgsb_minimize(bore_mlp *, int, int, const double *, int, const double *, const double *,
                         int, double, double, int, int, int, void *, size_t, double *, double *,
                         int32_t *, int32_t *, int32_t *, int32_t *, int *, long long *, void *) {
  bore_set_error("bore_lbfgsb_minimize: not implemented yet"); return -1; }
int bore_lbfgsb_init(const double *, int, int, const double *, const double *, int, double, double,
                     int, int, int, void *, size_t, double *, int32_t *, int, void *) {
  bore_set_error("bore_lbfgsb_init: not implemented yet"); return -1; }
int bore_lbfgsb_step(const void *, const void *, int, int, int, void *, double *, int32_t *, int *,
                     int, void *) {
  bore_set_error("bore_lbfgsb_step: not implemented yet"); return -1; }
int bore_lbfgsb_results(int, int, void *, double *, double *, int32_t *, int32_t *, int32_t *,
                        int32_t *, int, void *) {
  bore_set_error("bore_lbfgsb_results: not implemented yet"); return -1; }
int bore_topk_smallest(const float *, int, int, int32_t *, void *, size_t, int, void *) {
  bore_set_error("bore_topk_smallest: not implemented yet"); return -1; }
size_t bore_topk_workspace_bytes(int, int) { return 0; }
int bore_select_best(const double *, const int32_t *, const uint8_t *, int, int64_t, int64_t *, int,
                     void *) {
  bore_set_error("bore_select_best: not implemented yet"); return -1; }
}
