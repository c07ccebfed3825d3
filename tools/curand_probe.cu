// Probe (GPU box): facts about cuRAND the engine and oracle depend on.
//  1. does the HOST XORWOW generator reproduce the DEVICE generator's stream (same seed)?
//  2. is curandSetGeneratorOffset(k) == "start at element k" on the device generator (k multiple of 4096 / not)?
//  3. can curandGenerateNormal be captured into a CUDA graph, and does replay advance the stream?
//  4. throughput of curandGenerateNormal at the benchmark sizes.
#include <curand.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
#define CK(x) do{auto e=(x); if(e){printf("ERR %s:%d %d\n",__FILE__,__LINE__,(int)e);}}while(0)
int main(){
  const size_t n = 8192*100;
  curandGenerator_t gh, gd;
  CK(curandCreateGeneratorHost(&gh, CURAND_RNG_PSEUDO_DEFAULT));
  CK(curandCreateGenerator(&gd, CURAND_RNG_PSEUDO_DEFAULT));
  CK(curandSetPseudoRandomGeneratorSeed(gh, 42ULL)); CK(curandSetPseudoRandomGeneratorSeed(gd, 42ULL));
  std::vector<float> h(n), d(n), d2(n);
  float* dev; CK(cudaMalloc(&dev, n*4*8));
  CK(curandGenerateNormal(gh, h.data(), n, 0.f, 1.f));
  CK(curandGenerateNormal(gd, dev, n, 0.f, 1.f));
  CK(cudaMemcpy(d.data(), dev, n*4, cudaMemcpyDeviceToHost));
  size_t nd=0; for(size_t i=0;i<n;i++) nd += (h[i]!=d[i]);
  printf("[1] host==device stream: %s (mismatches %zu / %zu) first: %g %g | %g %g\n", nd?"NO":"YES", nd, n, h[0],h[1],d[0],d[1]);
  // second call continuity host vs device
  CK(curandGenerateNormal(gh, h.data(), n, 0.f, 1.f));
  CK(curandGenerateNormal(gd, dev, n, 0.f, 1.f));
  CK(cudaMemcpy(d2.data(), dev, n*4, cudaMemcpyDeviceToHost));
  nd=0; for(size_t i=0;i<n;i++) nd += (h[i]!=d2[i]);
  printf("[1b] second call host==device: %s (%zu)\n", nd?"NO":"YES", nd);
  // [2] offsets on device
  for (unsigned long long off : {409600ULL, 4096ULL, 8192ULL, 1000ULL, 2ULL}) {
    CK(curandSetGeneratorOffset(gd, off));
    CK(curandGenerateNormal(gd, dev, 4096, 0.f, 1.f));
    CK(cudaMemcpy(d2.data(), dev, 4096*4, cudaMemcpyDeviceToHost));
    printf("[2] device offset %llu == element index: %d\n", off, !memcmp(d2.data(), d.data()+off, 4096*4));
  }
  // [3] graph capture
  cudaStream_t s; CK(cudaStreamCreate(&s)); CK(curandSetStream(gd, s));
  CK(curandSetGeneratorOffset(gd, 0));
  CK(curandGenerateNormal(gd, dev, n, 0.f, 1.f)); // warm (allocs)
  CK(cudaStreamSynchronize(s));
  CK(curandSetGeneratorOffset(gd, 0));
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  curandStatus_t cs = curandGenerateNormal(gd, dev, n, 0.f, 1.f);
  cudaError_t e2 = cudaStreamEndCapture(s, &g);
  printf("[3] capture begin=%d curand=%d end=%d\n", (int)e, (int)cs, (int)e2);
  if (!e && !cs && !e2) {
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, s)); CK(cudaStreamSynchronize(s));
    CK(cudaMemcpy(d2.data(), dev, n*4, cudaMemcpyDeviceToHost));
    printf("[3] replay#1 == first block: %d\n", !memcmp(d2.data(), d.data(), n*4));
    CK(cudaGraphLaunch(ge, s)); CK(cudaStreamSynchronize(s));
    std::vector<float> d3(n); CK(cudaMemcpy(d3.data(), dev, n*4, cudaMemcpyDeviceToHost));
    printf("[3] replay#2 == first block again (stream NOT advanced): %d ; == second block (advanced): %d\n",
           !memcmp(d3.data(), d.data(), n*4), !memcmp(d3.data(), h.data(), n*4));
  } else { cudaGetLastError(); }
  // [4] throughput
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (size_t m : {(size_t)8192*100, (size_t)16384*150*2, (size_t)32768*100*2, (size_t)65536*150*2}) {
    if (m > n*8) continue;
    for(int i=0;i<3;i++) curandGenerateNormal(gd, dev, m, 0.f, 1.f);
    cudaEventRecord(a, s); for(int i=0;i<20;i++) curandGenerateNormal(gd, dev, m, 0.f, 1.f); cudaEventRecord(b, s);
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b);
    printf("[4] curandGenerateNormal n=%zu : %.2f us/call, %.1f GB/s written\n", m, ms*1000/20, m*4/(ms/20*1e-3)/1e9);
  }
  int v; curandGetVersion(&v); printf("curand version %d\n", v);
  return 0;
}
