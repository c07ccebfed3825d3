// Probe (GPU box): which texel does a point-sampled, normalised-coordinate tex2D return at exact texel boundaries?
// (reference golden 1116.3333 in tests/cost_functions/autorally_standard_cost_test.cu:958 implies texel-1)
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__global__ void k(cudaTextureObject_t t, const float* u, const float* v, float4* out, int n){
  int i = threadIdx.x; if (i<n) out[i] = tex2D<float4>(t, u[i], v[i]);
}
int main(){
  const int W=600,H=600; std::vector<float4> h(W*H);
  for(int i=0;i<H;i++)for(int j=0;j<W;j++) h[i*W+j]=make_float4((float)j,(float)i,0,0);
  cudaArray_t arr; auto ch=cudaCreateChannelDesc(32,32,32,32,cudaChannelFormatKindFloat);
  cudaMallocArray(&arr,&ch,W,H); cudaMemcpy2DToArray(arr,0,0,h.data(),W*16,W*16,H,cudaMemcpyHostToDevice);
  cudaResourceDesc r{}; r.resType=cudaResourceTypeArray; r.res.array.array=arr;
  cudaTextureDesc td{}; td.addressMode[0]=td.addressMode[1]=cudaAddressModeClamp; td.filterMode=cudaFilterModePoint; td.readMode=cudaReadModeElementType; td.normalizedCoords=1;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex,&r,&td,nullptr);
  std::vector<float> u,v;
  float r_c=1.0f/30.0f, tx=13.0f/30.0f, ty=10.0f/30.0f;
  float xs[]={3.0f,3.0f, 0.0f, 0.01f, 2.999f}; float ys[]={0.5f,-0.5f, 0.0f, 0.26f, 0.499f};
  for(int i=0;i<5;i++){ u.push_back(r_c*xs[i]+0.0f*ys[i]+tx); v.push_back(0.0f*xs[i]+r_c*ys[i]+ty);} 
  // sweep around a boundary: u*600 = 320 +- k ulps
  for(int k=-8;k<=8;k++){ float b=320.0f/600.0f; int ib; memcpy(&ib,&b,4); ib+=k; float f; memcpy(&f,&ib,4); u.push_back(f); v.push_back(0.5f/600.0f*401);} 
  for(int k=0;k<=16;k++){ u.push_back((320.0f + (k-8)/16.0f)/600.0f); v.push_back(0.5f);} 
  int n=u.size(); float *ud,*vd; float4* od; cudaMalloc(&ud,n*4);cudaMalloc(&vd,n*4);cudaMalloc(&od,n*16);
  cudaMemcpy(ud,u.data(),n*4,cudaMemcpyHostToDevice);cudaMemcpy(vd,v.data(),n*4,cudaMemcpyHostToDevice);
  k<<<1,64>>>(tex,ud,vd,od,n); std::vector<float4> o(n); cudaMemcpy(o.data(),od,n*16,cudaMemcpyDeviceToHost);
  for(int i=0;i<n;i++) printf("u=%.9g (u*W=%.6f) v=%.9g (v*H=%.6f) -> texel (%g,%g)\n",u[i],(double)u[i]*W,v[i],(double)v[i]*H,o[i].x,o[i].y);
  printf("err %d\n",(int)cudaGetLastError());
}
