// Build-container check (no GPU): runs the __host__ instantiation of the shuffle that csrc/datasplit.cu's kernel
// executes and prints the indices, for tests/test_datasplit_oracle.py to compare with oracle/datasplit_oracle.py.
//   nvcc -o oracle/_ref/check_datasplit_host oracle/check_datasplit_host.cu neural-process-family_b200/csrc/build/core.o
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../neural-process-family_b200/csrc/datasplit.cu"

int main(int argc, char** argv) {
    if (argc != 5) return 2;
    const int B = atoi(argv[1]), N = atoi(argv[2]), n = atoi(argv[3]);
    const unsigned long long seed = strtoull(argv[4], nullptr, 10);
    std::vector<int32_t> perm(N);
    for (int b = 0; b < B; ++b) {
        for (int t = 0; t < N; ++t) perm[t] = t;
        npf::partial_shuffle(perm.data(), N, n, (uint32_t)b, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int t = 0; t < n; ++t) printf("%d%c", perm[t], t + 1 == n ? '\n' : ' ');
    }
    return 0;
}
