#include <cstdio>
#include "../spatial-temporal-lidar-camera-calibration_b200/csrc/se3.cuh"
using namespace stl;
__device__ __noinline__ double he(const double* in, double* dbg) {
    const double *cR = in, *ct = in + 9; const double s = in[12]; const double *tl = in + 13; const double *tcd = in + 25;
    double TcR[9], Tct[3], TlR[9], Tlt[3], C1R[9], C1t[3], C2R[9], C2t[3], l1[6], l2[6];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { TcR[i * 3 + j] = tcd[i * 4 + j]; TlR[i * 3 + j] = tl[i * 4 + j]; }
        Tct[i] = stl::dmul(tcd[i * 4 + 3], s);
        Tlt[i] = tl[i * 4 + 3];
    }
    rt_compose(cR, ct, TlR, Tlt, C1R, C1t);
    rt_compose(TcR, Tct, cR, ct, C2R, C2t);
    se3_log(C1R, C1t, l1);
    se3_log(C2R, C2t, l2);
    for (int i=0;i<6;i++){dbg[i]=l1[i]; dbg[6+i]=l2[i];}
    for (int i=0;i<3;i++){dbg[12+i]=C1t[i]; dbg[15+i]=C2t[i];}
    double ss = 0;
    for (int i = 0; i < 6; ++i) { const double d = l1[i] - l2[i]; ss += d * d; }
    return sqrt(ss);
}
__global__ void k(const double* in, double* out) { out[0] = he(in, out+1); }
int main(){
  double in[37]; FILE* f=fopen("scripts/he_in.bin","rb"); fread(in,8,37,f); fclose(f);
  double *din,*dout; cudaMalloc(&din,37*8); cudaMalloc(&dout,8*32);
  cudaMemcpy(din,in,37*8,cudaMemcpyHostToDevice);
  k<<<1,1>>>(din,dout); double out[32]; cudaMemcpy(out,dout,8*32,cudaMemcpyDeviceToHost);
  for(int i=0;i<19;i++) printf("%.12g ", out[i]); printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
}
