// power_host.inl -- Matrix::power on the device (libmints/matrix.cc:2370-2424), the one O(naux^3) step of the DF setup
// (DFHelper::prepare_metric -> compute_metric, lib3index/dfhelper.cc:1462-1517: J^-1/2 for Ppq_, J^-1 for m1Ppq_).
// Included at the end of engine.cu.
//
//   eigendecomposition   cusolverDnDsyevd (a plain library call, resolved with dlopen like NCCL: the setup path, once
//                        per SCF, is not the hot path and the engine has no link-time dependency on it)
//   drop rule            exactly the reference's, on the host from the n eigenvalues (the same libm pow):
//                        alpha < 0 and |lambda| < cutoff * max|lambda|  -> 0;  non-finite lambda^alpha -> 0
//   V f(Lambda) V^T      scale kernel + the strided DMMA GEMM (C_DSCAL + C_DGEMM('T','N') at :2412-2417)
//
// At naux = 4740 the reference spends tens of seconds in LAPACK here on the host cores.

namespace {

typedef struct cusolverDnContext* cusolverDnHandle_t;
struct CuSolver {
    void* lib = nullptr;
    int (*Create)(cusolverDnHandle_t*) = nullptr;
    int (*Destroy)(cusolverDnHandle_t) = nullptr;
    int (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    int (*Dsyevd_bufferSize)(cusolverDnHandle_t, int, int, int, const double*, int, const double*, int*) = nullptr;
    int (*Dsyevd)(cusolverDnHandle_t, int, int, int, double*, int, double*, double*, int, int*) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so", nullptr};
        for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            err = std::string("cannot dlopen libcusolver: ") + dlerror();
            return false;
        }
#define SYM(f)                                          \
    *(void**)(&f) = dlsym(lib, "cusolverDn" #f);        \
    if (!f) {                                           \
        err = "libcusolver lacks cusolverDn" #f;        \
        return false;                                   \
    }
        SYM(Create) SYM(Destroy) SYM(SetStream) SYM(Dsyevd_bufferSize) SYM(Dsyevd)
#undef SYM
        return true;
    }
};
CuSolver g_cusolver;

// rows of the eigenvector matrix scaled by f(lambda_i): A2[i][:] = s[i] * A1[i][:]
__global__ void power_scale_rows_kernel(const double* __restrict__ A1, const double* __restrict__ s, size_t n,
                                        double* __restrict__ A2) {
    const size_t i = blockIdx.y;
    const double f = s[i];
    for (size_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
        A2[i * n + c] = f * A1[i * n + c];
}

}  // namespace

extern "C" int b200jk_matrix_power(b200jk_t* h, size_t n, const double* A, double alpha, double cutoff, double* out,
                                   int* remaining, double* ms_device) {
    if (!h) return B200JK_ERR_INVALID;
    if (h->sh.empty()) return fail(h, B200JK_ERR_NODEVICE, "handle has no device");
    if (!n || n > 46000 || !A || !out) return fail(h, B200JK_ERR_INVALID, "matrix_power: bad arguments");
    if (!g_cusolver.load(h->err)) return B200JK_ERR_CUDA;
    Shard& s = h->sh[0];
    CK(cudaSetDevice(s.dev));
    const size_t n2 = n * n;
    double *dA = nullptr, *dA2 = nullptr, *dW = nullptr, *dwork = nullptr, *dS = nullptr, *dOut = nullptr;
    int* dinfo = nullptr;
    cusolverDnHandle_t cs = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        if (cs) g_cusolver.Destroy(cs);
        void* ptrs[] = {dA, dA2, dW, dwork, dS, dOut, dinfo};
        for (void* p : ptrs)
            if (p) cudaFree(p);
    };
#define PK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            rc = fail(h, e_ == cudaErrorMemoryAllocation ? B200JK_ERR_OOM : B200JK_ERR_CUDA, "%s: %s", #call, \
                      cudaGetErrorString(e_));                                                        \
            cleanup();                                                                                \
            return rc;                                                                                \
        }                                                                                             \
    } while (0)
    PK(cudaMalloc((void**)&dA, n2 * 8));
    PK(cudaMalloc((void**)&dA2, n2 * 8));
    PK(cudaMalloc((void**)&dOut, n2 * 8));
    PK(cudaMalloc((void**)&dW, n * 8));
    PK(cudaMalloc((void**)&dS, n * 8));
    PK(cudaMalloc((void**)&dinfo, sizeof(int)));
    PK(cudaMemcpyAsync(dA, A, n2 * 8, cudaMemcpyHostToDevice, s.stream));
    cudaEvent_t e0, e1;
    PK(cudaEventCreate(&e0));
    PK(cudaEventCreate(&e1));
    PK(cudaEventRecord(e0, s.stream));
    int lwork = 0;
    // row-major symmetric input == column-major symmetric input; the eigenvectors come back as the COLUMNS of the
    // column-major matrix, i.e. as the ROWS of the row-major view -- the layout C_DSYEV leaves in A1 (matrix.cc:2391)
    if (g_cusolver.Create(&cs) || g_cusolver.SetStream(cs, s.stream) ||
        g_cusolver.Dsyevd_bufferSize(cs, 1 /*CUSOLVER_EIG_MODE_VECTOR*/, 1 /*CUBLAS_FILL_MODE_UPPER*/, (int)n, dA, (int)n, dW,
                                     &lwork)) {
        cleanup();
        return fail(h, B200JK_ERR_CUDA, "matrix_power: cusolverDnDsyevd_bufferSize failed");
    }
    PK(cudaMalloc((void**)&dwork, (size_t)std::max(lwork, 1) * 8));
    if (g_cusolver.Dsyevd(cs, 1, 1, (int)n, dA, (int)n, dW, dwork, lwork, dinfo)) {
        cleanup();
        return fail(h, B200JK_ERR_CUDA, "matrix_power: cusolverDnDsyevd failed");
    }
    std::vector<double> w(n), sc(n);
    int info = 0;
    PK(cudaMemcpyAsync(w.data(), dW, n * 8, cudaMemcpyDeviceToHost, s.stream));
    PK(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    PK(cudaStreamSynchronize(s.stream));
    if (info) {
        cleanup();
        return fail(h, B200JK_ERR_CUDA, "Matrix::power: eigendecomposition failed (info %d)", info);  // matrix.cc:2396
    }
    // matrix.cc:2399-2411
    const double max_a = std::fabs(w[n - 1]) > std::fabs(w[0]) ? std::fabs(w[n - 1]) : std::fabs(w[0]);
    int remain = 0;
    for (size_t i = 0; i < n; i++) {
        double a = w[i];
        if (alpha < 0.0 && std::fabs(a) < cutoff * max_a) {
            a = 0.0;
        } else {
            a = pow(a, alpha);
            if (std::isfinite(a))
                remain++;
            else
                a = 0.0;
        }
        sc[i] = a;
    }
    PK(cudaMemcpyAsync(dS, sc.data(), n * 8, cudaMemcpyHostToDevice, s.stream));
    power_scale_rows_kernel<<<dim3((unsigned)std::min<size_t>((n + 255) / 256, 64), (unsigned)n), 256, 0, s.stream>>>(dA, dS, n, dA2);
    PK(cudaGetLastError());
    // out[r][c] = sum_i A2[i][r] A1[i][c]   (C_DGEMM('T','N'), matrix.cc:2417)
    {
        int grc = launch_gemm(h, s, gemm_desc((int)n, (int)n, (int)n, dA2, 1, (long long)n, dA, 1, (long long)n, dOut,
                                              (long long)n, 1, 1.0, 0.0));
        if (grc) {
            cleanup();
            return grc;
        }
    }
    PK(cudaEventRecord(e1, s.stream));
    PK(cudaMemcpyAsync(out, dOut, n2 * 8, cudaMemcpyDeviceToHost, s.stream));
    PK(cudaStreamSynchronize(s.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_device) *ms_device = ms;
    if (remaining) *remaining = remain;
    cleanup();
#undef PK
    return 0;
}
