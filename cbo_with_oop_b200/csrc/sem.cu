// K6 -- ground truth of an intervention: E[target | do(X = x)] of a structural equation model by Monte Carlo, on the device.
//
// Replaces compute_interventions / sample_from_model / intervene_dict (reference graph_functions.py:8-77): the reference
// draws 100 000 samples of the mutilated SEM in a Python loop per intervention -- the slowest thing in a runCBO.py run
// (SURVEY.md §3, §8f.3).  compute_interventions reseeds with seed 1 on every call (graph_functions.py:73), so the noise is a
// CONSTANT of the run: the host draws it once with the same NumPy stream (bit-identical), it stays resident in HBM, and a
// batch of interventions is one launch.
//
// The SEM is a small program (cbo_sem_node / cbo_sem_term): node values in topological order,
//     value[node] = constant + sum_t coef_t * f_t(scale_t * value[src_t]) ,   f in {x, exp, cos, sin, x^2}
// with sources either noise columns (src < num_noise) or earlier nodes (src - num_noise); an intervened node takes the
// intervention's constant instead (the mutilated model).  One thread = one sample at a time, all node values in registers;
// per-block partial sums in a fixed order, then one block per intervention adds them (deterministic).
// Roofline: HBM -- 8 B x num_noise per sample and intervention (the noise matrix is L2 resident across a batch);
// transcendental-bound in practice for the shipped graphs (a handful of exp / cos per sample).
#include "cbo_common.cuh"

namespace cbo {

constexpr int kSemBlocks = CBO_SEM_BLOCKS;
constexpr int kSemThreads = 256;

__global__ void __launch_bounds__(kSemThreads)
sem_eval_kernel(const cbo_sem_node* __restrict__ nodes, int num_nodes, const cbo_sem_term* __restrict__ terms, int num_terms,
                const double* __restrict__ noise, int num_noise, long long num_samples, const int32_t* __restrict__ do_mask,
                const double* __restrict__ do_value, int target_node, double* __restrict__ partials) {
    __shared__ cbo_sem_node s_nodes[CBO_SEM_MAX_NODES];
    __shared__ cbo_sem_term s_terms[CBO_SEM_MAX_TERMS];
    __shared__ int s_mask[CBO_SEM_MAX_NODES];
    __shared__ double s_val[CBO_SEM_MAX_NODES];
    __shared__ double red[kSemThreads / 32];
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < num_nodes; i += kSemThreads) {
        s_nodes[i] = nodes[i];
        s_mask[i] = do_mask[(size_t)b * num_nodes + i];
        s_val[i] = do_value[(size_t)b * num_nodes + i];
    }
    for (int i = tid; i < num_terms; i += kSemThreads) s_terms[i] = terms[i];
    __syncthreads();
    double sum = 0.0;
    for (long long n = (long long)blockIdx.x * kSemThreads + tid; n < num_samples; n += (long long)gridDim.x * kSemThreads) {
        double v[CBO_SEM_MAX_NODES];
#pragma unroll
        for (int i = 0; i < CBO_SEM_MAX_NODES; ++i) {
            if (i < num_nodes) {
                double x = s_nodes[i].constant;
                const int t0 = s_nodes[i].first_term, t1 = t0 + s_nodes[i].num_terms;
                for (int t = t0; t < t1; ++t) {
                    const cbo_sem_term T = s_terms[t];
                    double a;
                    if (T.src < num_noise) a = noise[(size_t)T.src * num_samples + n];
                    else {
                        // earlier node: a dynamic index into registers would go through local memory; nodes are few, so select
                        a = 0.0;
#pragma unroll
                        for (int j = 0; j < CBO_SEM_MAX_NODES; ++j)
                            if (j == T.src - num_noise) a = v[j];
                    }
                    a *= T.scale;
                    double f;
                    switch (T.func) {
                        case CBO_SEM_EXP: f = exp(a); break;
                        case CBO_SEM_COS: f = cos(a); break;
                        case CBO_SEM_SIN: f = sin(a); break;
                        case CBO_SEM_SQUARE: f = a * a; break;
                        default: f = a; break;
                    }
                    x += T.coef * f;
                }
                v[i] = s_mask[i] ? s_val[i] : x;
            } else {
                v[i] = 0.0;
            }
        }
        double y = 0.0;
#pragma unroll
        for (int j = 0; j < CBO_SEM_MAX_NODES; ++j)
            if (j == target_node) y = v[j];
        sum += y;
    }
    sum = warp_sum(sum);
    if ((tid & 31) == 0) red[tid >> 5] = sum;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int x = 0; x < kSemThreads / 32; ++x) t += red[x];
        partials[(size_t)b * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void sem_mean_kernel(const double* __restrict__ partials, int blocks, int batch, long long num_samples,
                                double* __restrict__ mean) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double t = 0.0;
    for (int i = 0; i < blocks; ++i) t += partials[(size_t)b * blocks + i];
    mean[b] = t / (double)num_samples;
}

int sem_eval_impl(const cbo_sem_node* d_nodes, int num_nodes, const cbo_sem_term* d_terms, int num_terms, const double* d_noise,
                  int num_noise, long long num_samples, const int32_t* d_do_mask, const double* d_do_value, int batch,
                  int target_node, double* d_partials, double* d_mean, cudaStream_t st) {
    CBO_REQUIRE(d_nodes && d_terms && d_do_mask && d_do_value && d_partials && d_mean, "cbo_sem_eval: NULL pointer");
    CBO_REQUIRE(num_nodes >= 1 && num_nodes <= CBO_SEM_MAX_NODES, "cbo_sem_eval: num_nodes=%d outside [1,%d]", num_nodes, CBO_SEM_MAX_NODES);
    CBO_REQUIRE(num_terms >= 0 && num_terms <= CBO_SEM_MAX_TERMS, "cbo_sem_eval: num_terms=%d outside [0,%d]", num_terms, CBO_SEM_MAX_TERMS);
    CBO_REQUIRE(num_noise >= 0 && (num_noise == 0 || d_noise), "cbo_sem_eval: noise matrix missing");
    CBO_REQUIRE(num_samples >= 1 && batch >= 1 && batch <= 65535, "cbo_sem_eval: num_samples=%lld batch=%d", num_samples, batch);
    CBO_REQUIRE(target_node >= 0 && target_node < num_nodes, "cbo_sem_eval: target_node=%d outside the program", target_node);
    sem_eval_kernel<<<dim3(kSemBlocks, (unsigned)batch), kSemThreads, 0, st>>>(d_nodes, num_nodes, d_terms, num_terms, d_noise, num_noise,
                                                                               num_samples, d_do_mask, d_do_value, target_node, d_partials);
    note_launch();
    CBO_CUDA(cudaGetLastError());
    sem_mean_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, st>>>(d_partials, kSemBlocks, batch, num_samples, d_mean);
    note_launch();
    CBO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cbo
