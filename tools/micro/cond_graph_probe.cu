#include <cuda_runtime.h>
#include <stdio.h>
__global__ void set_cond(cudaGraphConditionalHandle h, const int* flag) { if (threadIdx.x == 0) cudaGraphSetConditional(h, *flag != 0); }
__global__ void body(int* out) { atomicAdd(out, 1); }
int main() {
    int *d_flag, *d_out; cudaMalloc(&d_flag, 4); cudaMalloc(&d_out, 4); cudaMemset(d_out, 0, 4);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
    cudaGraphNode_t n0; cudaKernelNodeParams kp = {}; void* args[] = {&h, &d_flag};
    kp.func = (void*)set_cond; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.kernelParams = args;
    printf("add kernel %d\n", cudaGraphAddKernelNode(&n0, g, nullptr, 0, &kp));
    cudaGraphNodeParams cp = {}; cp.type = cudaGraphNodeTypeConditional; cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeIf; cp.conditional.size = 1;
    cudaGraphNode_t nc; printf("add cond %d\n", cudaGraphAddNode(&nc, g, &n0, 1, &cp));
    cudaGraph_t bodyg = cp.conditional.phGraph_out[0];
    cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
    body<<<1, 1, 0, s>>>(d_out); body<<<1, 1, 0, s>>>(d_out);
    printf("end capture %d\n", cudaStreamEndCapture(s, nullptr));
    cudaGraphExec_t ge; printf("instantiate %d\n", cudaGraphInstantiate(&ge, g, 0));
    for (int f = 0; f < 2; ++f) { cudaMemcpy(d_flag, &f, 4, cudaMemcpyHostToDevice); cudaGraphLaunch(ge, s); cudaStreamSynchronize(s); int o; cudaMemcpy(&o, d_out, 4, cudaMemcpyDeviceToHost); printf("flag %d -> out %d (%s)\n", f, o, cudaGetErrorString(cudaGetLastError())); }
    return 0;
}
