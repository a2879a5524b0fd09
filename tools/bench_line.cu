// Micro-benchmark of the hybrid line smoother kernel (phase time stamps of block 0, whole-launch times per level size).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DTPB_LINE_TIMING -Iinclude -Ithermalporous_b200/csrc \
//        tools/bench_line.cu thermalporous_b200/_obj/{tpb_api,tpb_assemble,tpb_blas,tpb_comm,tpb_solver,tpb_spmv}.o -ldl -o tools/bench_line
#include "../thermalporous_b200/csrc/tpb_pc.cu"

#include <vector>

int main() {
    const int nz = 85;
    const int shapes[5][2] = {{60, 220}, {60, 110}, {30, 55}, {15, 28}, {8, 14}};
    cudaStream_t st;
    cudaStreamCreate(&st);
    for (int q = 0; q < 5; q++) {
        const int nx = shapes[q][0], ny = shapes[q][1];
        const long long n = (long long)nx * ny * nz;
        std::vector<double> ha(7 * n), hb(n);
        for (long long c = 0; c < n; c++) {
            ha[c] = 6.5;
            for (int s = 1; s < 7; s++) ha[s * n + c] = -1.0;
            hb[c] = (double)(c % 17) - 8.0;
        }
        double *a, *fac, *b, *x0, *x1;
        cudaMalloc(&a, 7 * n * 8); cudaMalloc(&fac, 3 * n * 8); cudaMalloc(&b, n * 8); cudaMalloc(&x0, n * 8); cudaMalloc(&x1, n * 8);
        cudaMemcpy(a, ha.data(), 7 * n * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(b, hb.data(), n * 8, cudaMemcpyHostToDevice);
        cudaMemset(x0, 0, n * 8);
        LevGeom g{nx, ny, nz, 2, 2, 1, n, 1};
        line_factor_kernel<<<(nx * ny + 127) / 128, 128, 0, st>>>(a, g, fac);
        int tx, ty;
        line_tile_shape(nz, tx, ty);
        const unsigned grid = ((nx + tx - 1) / tx) * ((ny + ty - 1) / ty);
        const size_t smem = line_smem_doubles(nz, tx, ty) * 8;
        cudaFuncSetAttribute(line_smooth_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int nsw = 1; nsw <= 2; nsw++) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int w = 0; w < 3; w++)
                line_smooth_kernel<false><<<grid, LS_THREADS, smem, st>>>(a, fac, b, x0, x1, g, nullptr, 0, 0, 0.0, nsw, tx, ty);
            cudaEventRecord(e0, st);
            const int reps = 20;
            for (int r = 0; r < reps; r++)
                line_smooth_kernel<false><<<grid, LS_THREADS, smem, st>>>(a, fac, b, x0, x1, g, nullptr, 0, 0, 0.0, nsw, tx, ty);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            long long clk[8];
            cudaMemcpyFromSymbol(clk, g_line_clk, sizeof(clk));
            printf("%3dx%3dx%d grid %4u nsw %d: %.2f us/launch (back to back) | block 0 cycles: statics %lld, x+sync %lld, rhs %lld, solve %lld, "
                   "rest of passes %lld, store %lld, total %lld  [%s]\n", nx, ny, nz, grid, nsw, ms * 1e3 / reps, clk[1] - clk[0], clk[2] - clk[1],
                   clk[3] - clk[2], clk[4] - clk[3], clk[5] - clk[4], clk[6] - clk[5], clk[6] - clk[0], cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(a); cudaFree(fac); cudaFree(b); cudaFree(x0); cudaFree(x1);
    }
    return 0;
}
