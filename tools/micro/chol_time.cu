#include <cstdio>
#include <vector>
#include <cmath>
#define TVO_CHOL_TIMING
#include "bsc_chol.cuh"
using namespace tvo;
int main(){ int H=256,D=144; std::vector<double> st((size_t)H*D+H*H+H+3,0.0);
 for(int i=0;i<H;i++) for(int j=i;j<H;j++) st[(size_t)H*D+i*H+j]= (i==j)? 10.0+ i*0.01 : 0.01*std::sin(i*0.3+j*0.7);
 for(int i=0;i<H*D;i++) st[i]=std::cos(i*0.1);
 double *ds,*work,*sol,*fl; cudaMalloc(&ds,st.size()*8); cudaMemcpy(ds,st.data(),st.size()*8,cudaMemcpyHostToDevice);
 cudaMalloc(&work,(chol_work_doubles(H)+2)*8); cudaMalloc(&sol,H*D*8); fl=work+chol_work_doubles(H);
 size_t smem=chol_factor_smem(H); cudaFuncSetAttribute(bsc_chol_factor_kernel,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
 size_t sm2=chol_solve_smem(H); cudaFuncSetAttribute(bsc_chol_solve_kernel,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm2);
 size_t smc=chol_cluster_smem(H); cudaFuncSetAttribute(bsc_chol_factor_cluster_kernel,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smc);
 cudaEvent_t e0,e1,e2; cudaEventCreate(&e0);cudaEventCreate(&e1);cudaEventCreate(&e2);
 for(int it=0;it<3;it++){ cudaEventRecord(e0); bsc_chol_factor_kernel<<<1,256,smem>>>(ds+(size_t)H*D,H,work,fl); cudaEventRecord(e1);
  bsc_chol_solve_kernel<<<D,256,sm2>>>(work,ds,sol,fl,H,D); cudaEventRecord(e2); cudaDeviceSynchronize();
  float a,b; cudaEventElapsedTime(&a,e0,e1); cudaEventElapsedTime(&b,e1,e2); printf("factor %.1f us solve %.1f us err %s\n",a*1e3,b*1e3,cudaGetErrorString(cudaGetLastError()));
  cudaEventRecord(e0); bsc_chol_factor_cluster_kernel<<<kChCl,256,smc>>>(ds+(size_t)H*D,H,work,fl); cudaEventRecord(e1); cudaDeviceSynchronize();
  cudaEventElapsedTime(&a,e0,e1); printf("cluster factor %.1f us err %s\n",a*1e3,cudaGetErrorString(cudaGetLastError())); }
 double f[2]; cudaMemcpy(f,fl,16,cudaMemcpyDeviceToHost); printf("flags %g %g\n",f[0],f[1]); return 0; }
