// micro-benchmark: launch overhead of (a) stream launches, (b) plain graph, (c) graph WHILE node, (d) IF nodes
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(256) sync_kernel(int* p, int n) { cg::grid_group g = cg::this_grid(); for (int i = 0; i < n; ++i) { if (threadIdx.x == 0 && blockIdx.x == (i % gridDim.x)) p[0] += 1; g.sync(); } }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 1000) *p = 1; }
__global__ void work_kernel(int* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] += 1; }
__global__ void loop_ctl(cudaGraphConditionalHandle h, int* counter, int iters) {
  if (threadIdx.x == 0) { int c = ++*counter; cudaGraphSetConditional(h, c < iters ? 1 : 0); }
}
__global__ void if_ctl(cudaGraphConditionalHandle h, int v) { if (threadIdx.x == 0) cudaGraphSetConditional(h, v); }
__global__ void reset(int* c) { *c = 0; }

int main() {
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  int* d; CK(cudaMalloc(&d, 4 << 20)); CK(cudaMemset(d, 0, 4 << 20));
  int* cnt; CK(cudaMalloc(&cnt, 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int IT = 20, KPI = 5, REP = 20;
  float ms;
  // (a) stream launches of empty kernels
  for (int w = 0; w < 3; ++w) {
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < REP; ++r) for (int i = 0; i < IT * KPI; ++i) empty_kernel<<<296, 256, 0, s>>>(d);
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  printf("stream: %d empty launches: %.2f us each\n", IT * KPI, ms * 1e3 / (REP * IT * KPI));
  // (b) plain graph
  {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < IT * KPI; ++i) empty_kernel<<<296, 256, 0, s>>>(d);
    CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
    for (int w = 0; w < 3; ++w) {
      CK(cudaEventRecord(e0, s));
      for (int r = 0; r < REP; ++r) CK(cudaGraphLaunch(ge, s));
      CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("graph: %d empty kernel nodes: %.2f us each\n", IT * KPI, ms * 1e3 / (REP * IT * KPI));
  }
  // (c) WHILE node: body = KPI-1 empty kernels + loop_ctl
  for (int body = 1; body <= 5; body += 2) {
    cudaGraph_t g; cudaGraphExec_t ge; CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNode_t rn;
    { cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeKernel; void* args[] = {&cnt};
      p.kernel.func = (void*) reset; p.kernel.gridDim = dim3(1); p.kernel.blockDim = dim3(1); p.kernel.kernelParams = args;
      CK(cudaGraphAddNode(&rn, g, nullptr, 0, &p)); }
    cudaGraphNodeParams cp = {}; cp.type = cudaGraphNodeTypeConditional; cp.conditional.handle = h;
    cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
    cudaGraphNode_t cn; CK(cudaGraphAddNode(&cn, g, &rn, 1, &cp));
    cudaGraph_t bg = cp.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < body - 1; ++i) empty_kernel<<<296, 256, 0, s>>>(d);
    loop_ctl<<<1, 32, 0, s>>>(h, cnt, IT);
    CK(cudaStreamEndCapture(s, nullptr));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int w = 0; w < 3; ++w) {
      CK(cudaEventRecord(e0, s));
      for (int r = 0; r < REP; ++r) CK(cudaGraphLaunch(ge, s));
      CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    int hc; CK(cudaMemcpy(&hc, cnt, 4, cudaMemcpyDeviceToHost));
    printf("while graph: body of %d kernels x %d iterations (counter %d): %.2f us per iteration\n", body, IT, hc, ms * 1e3 / (REP * IT));
  }
  // (d) IF nodes inside a flat graph: per iteration [ctl kernel, IF{2 empty kernels}] with the condition false / true
  for (int v = 0; v <= 1; ++v) {
    cudaGraph_t g; cudaGraphExec_t ge; CK(cudaGraphCreate(&g, 0));
    cudaGraphNode_t prev; bool have = false;
    for (int it = 0; it < IT; ++it) {
      cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
      cudaGraphNode_t kn;
      { cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeKernel; void* args[] = {&h, &v};
        p.kernel.func = (void*) if_ctl; p.kernel.gridDim = dim3(296); p.kernel.blockDim = dim3(256); p.kernel.kernelParams = args;
        CK(cudaGraphAddNode(&kn, g, have ? &prev : nullptr, have ? 1 : 0, &p)); }
      cudaGraphNodeParams cp = {}; cp.type = cudaGraphNodeTypeConditional; cp.conditional.handle = h;
      cp.conditional.type = cudaGraphCondTypeIf; cp.conditional.size = 1;
      cudaGraphNode_t cn; CK(cudaGraphAddNode(&cn, g, &kn, 1, &cp));
      cudaGraph_t bg = cp.conditional.phGraph_out[0];
      CK(cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
      empty_kernel<<<296, 256, 0, s>>>(d); empty_kernel<<<296, 256, 0, s>>>(d);
      CK(cudaStreamEndCapture(s, nullptr));
      prev = cn; have = true;
    }
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int w = 0; w < 3; ++w) {
      CK(cudaEventRecord(e0, s));
      for (int r = 0; r < REP; ++r) CK(cudaGraphLaunch(ge, s));
      CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("if graph (cond=%d): [ctl kernel + IF{2 kernels}] : %.2f us per iteration\n", v, ms * 1e3 / (REP * IT));
  }
  // (e) cooperative grid sync cost
  for (int bpsm = 1; bpsm <= 2; ++bpsm) {
    int n0 = 0, n1 = 200; float t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      int n = pass ? n1 : n0; void* args[] = {&d, &n};
      for (int w = 0; w < 3; ++w) {
        CK(cudaEventRecord(e0, s));
        CK(cudaLaunchCooperativeKernel((void*) sync_kernel, dim3(148 * bpsm), dim3(256), args, 0, s));
        CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      (pass ? t1 : t0) = ms;
    }
    printf("cooperative: %d blocks: launch %.2f us, grid.sync %.2f us each\n", 148 * bpsm, t0 * 1e3, (t1 - t0) * 1e3 / n1);
  }
  return 0;
}
