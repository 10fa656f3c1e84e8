#include <cuda_runtime.h>
#include <cstdio>
__global__ void body_kernel(int* counter) { atomicAdd(counter, 1); }
__global__ void cond_kernel(cudaGraphConditionalHandle h, int* counter, int limit) {
    cudaGraphSetConditional(h, *counter < limit ? 1u : 0u);
}
int main() {
    int* c; cudaMalloc(&c, 4); cudaMemset(c, 0, 4);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("add node: %s\n", cudaGetErrorString(e));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    e = cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    printf("begin capture: %s\n", cudaGetErrorString(e));
    body_kernel<<<1, 1, 0, s>>>(c);
    cond_kernel<<<1, 1, 0, s>>>(h, c, 7);
    e = cudaStreamEndCapture(s, nullptr);
    printf("end capture: %s\n", cudaGetErrorString(e));
    cudaGraphExec_t x; e = cudaGraphInstantiate(&x, g, 0);
    printf("instantiate: %s\n", cudaGetErrorString(e));
    cudaGraphLaunch(x, s); cudaStreamSynchronize(s);
    int hc; cudaMemcpy(&hc, c, 4, cudaMemcpyDeviceToHost);
    printf("counter %d (want 7) %s\n", hc, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
