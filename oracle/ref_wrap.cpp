// ref_wrap.cpp -- C-ABI wrappers around the REFERENCE's own serial CPU helpers
// (/root/reference/include/ICP/tests/helper_funcs.hpp), compiled where they lie (nothing is copied
// into this repo).  TEST INFRASTRUCTURE ONLY: used to pin oracle/icp_oracle.cpp and, optionally, as
// the "reference" CPU baseline.  Built by oracle/Makefile into oracle/_ref/libicp_ref.so.
//
// The helper header needs <CL/cl.hpp> (scalar typedefs only) and <RBC/data_types.hpp> (rbc_dist_id):
// oracle/ref_shim/ provides both.  cpuICPPowerMethod is declared `T ...` but never returns
// (helper_funcs.hpp:681-764) => g++ >= 8 plants a trap at its end; the Makefile therefore pipes the
// PREPROCESSED translation unit through `sed 's/T cpuICPPowerMethod/void cpuICPPowerMethod/'`
// before compiling it (no patched copy of the reference is written anywhere).
#include <ICP/tests/helper_funcs.hpp>

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void ref_ICPLMs(float *in, float *out) { ICP::cpuICPLMs<float>(in, out); }
REF_API void ref_ICPReps(float *in, float *out, uint32_t nr) { ICP::cpuICPReps<float>(in, out, nr); }
REF_API void ref_ICPWeights(rbc_dist_id *D, float *W, double *SW, uint32_t n) { ICP::cpuICPWeights<float>(D, W, SW, n); }
REF_API void ref_ICPMean(float *F, float *M, float *mean, uint32_t n) { ICP::cpuICPMean<float>(F, M, mean, n); }
REF_API void ref_ICPMeanWeighted(float *F, float *M, float *MEAN, float *W, uint32_t n) { ICP::cpuICPMeanWeighted<float>(F, M, MEAN, W, n); }
REF_API void ref_ICPDevs(float *F, float *M, float *DF, float *DM, float *mean, uint32_t n) { ICP::cpuICPDevs<float>(F, M, DF, DM, mean, n); }
REF_API void ref_ICPS(float *DM, float *DF, float *S, uint32_t m, float c) { ICP::cpuICPS<float>(DM, DF, S, m, c); }
REF_API void ref_ICPSw(float *M, float *F, float *W, float *S, uint32_t m, float c) { ICP::cpuICPSw<float>(M, F, W, S, m, c); }
REF_API void ref_ICPTransformQ(float *M, float *tM, float *D, uint32_t m) { ICP::cpuICPTransformQ<float>(M, tM, D, m); }
REF_API void ref_ICPTransformQ2(float *M, float *tM, float *D, uint32_t m) { ICP::cpuICPTransformQ2<float>(M, tM, D, m); }
REF_API void ref_ICPTransformM(float *M, float *tM, float *D, uint32_t m) { ICP::cpuICPTransformM<float>(M, tM, D, m); }
REF_API void ref_ICPPowerMethod(float *Sij, float *means, float *Tk) { ICP::cpuICPPowerMethod<float>(Sij, means, Tk); }
REF_API void ref_ReduceSum(float *in, float *out, uint32_t cols, uint32_t rows) { ICP::cpuReduceSum<float>(in, out, cols, rows); }
REF_API void ref_ReduceMin(float *in, float *out, uint32_t cols, uint32_t rows)
{ ICP::cpuReduce<float>(in, out, cols, rows, [](float a, float b) { return a < b; }); }
REF_API void ref_ReduceMaxU(uint32_t *in, uint32_t *out, uint32_t cols, uint32_t rows)
{ ICP::cpuReduce<uint32_t>(in, out, cols, rows, [](uint32_t a, uint32_t b) { return a > b; }); }
REF_API void ref_InScan(int *in, int *out, uint32_t w, uint32_t h) { ICP::cpuInScan<int>(in, out, w, h); }
REF_API void ref_ExScan(int *in, int *out, uint32_t w, uint32_t h) { ICP::cpuExScan<int>(in, out, w, h); }
