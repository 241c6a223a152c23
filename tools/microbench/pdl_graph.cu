// Does programmatic dependent launch survive stream capture into a CUDA graph, and what does it buy?
// Chain of 4 dependent kernels (each ~10 us of spinning on 148 CTAs + a prologue), timed: eager, eager + PDL,
// graph, graph + PDL; prints the edge types of the captured graph.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void work(float* buf, int iters, int pdl) {
  float x = threadIdx.x * 1e-3f;
  for (int i = 0; i < 200; ++i) x = fmaf(x, 1.0001f, 0.5f);          // "prologue" that needs no memory
  if (pdl) { asm volatile("griddepcontrol.wait;" ::"f"(x) : "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
  float v = buf[blockIdx.x * blockDim.x + threadIdx.x];
  for (int i = 0; i < iters; ++i) v = fmaf(v, 1.0001f, x);
  buf[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
static cudaError_t launch(float* buf, int iters, int pdl, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148 * 4); cfg.blockDim = dim3(256); cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, work, buf, iters, pdl);
}
int main() {
  float* buf; cudaMalloc(&buf, 148 * 4 * 256 * 4); cudaMemset(buf, 0, 148 * 4 * 256 * 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 3000, chain = 4, reps = 200;
  for (int pdl = 0; pdl < 2; ++pdl) {
    for (int i = 0; i < 20; ++i) launch(buf, iters, pdl, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; ++r) for (int i = 0; i < chain; ++i) launch(buf, iters, pdl, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("eager  pdl=%d: %.2f us per 4-kernel chain\n", pdl, ms * 1000 / reps);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < chain; ++i) launch(buf, iters, pdl, st);
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); continue; }
    size_t ne = 0; cudaGraphGetEdges_v2(g, nullptr, nullptr, nullptr, &ne);
    cudaGraphNode_t from[16], to[16]; cudaGraphEdgeData ed[16]; if (ne > 16) ne = 16;
    cudaGraphGetEdges_v2(g, from, to, ed, &ne);
    printf("graph  pdl=%d: %zu edges, types:", pdl, ne);
    for (size_t i = 0; i < ne; ++i) printf(" %d(from_port=%d)", (int)ed[i].type, (int)ed[i].from_port);
    printf("\n");
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); continue; }
    for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("graph  pdl=%d: %.2f us per 4-kernel chain (replay)\n", pdl, ms * 1000 / reps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
