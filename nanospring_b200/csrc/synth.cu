// Synthetic nanopore-like reads for the benchmark and the tests.  Follows the
// recipe of the reference's util/old_code/createData.py:
//   createRandomGenome (:67-76)   uniform i.i.d. bases over {A,T,C,G}
//   generateRead       (:88-148)  uniform start, wrap-around, per step an insertion
//                                 (p_ins), deletion (p_del), substitution (p_sub) or copy
//   generateData       (:196-201) half of the reads reverse-complemented
// but with a counter-based generator (splitmix64 finaliser), so the same read comes
// out of the host function and the device kernel, for any subset of reads, and the
// genome never has to be stored: base p of the genome is a pure function of p.
#include "nsmh_internal.cuh"

namespace nsmh {

struct SynthConst {
    uint64_t genome_len, genome_seed, read_seed;
    uint64_t t_ins, t_del, t_sub, t_rc;   // cumulative thresholds scaled to 2^64
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint32_t genome_code(const SynthConst &c, uint64_t p) {
    return (uint32_t)(mix64(c.genome_seed + (p + 1) * 0x9E3779B97F4A7C15ULL) >> 62);
}

// Writes read `rid` (length len) to out[0..len).
__host__ __device__ inline void synth_one_read(const SynthConst &c, uint64_t rid, uint64_t len, char *out) {
    const char letters[4] = {'A', 'T', 'C', 'G'};   // code order of baseToInt (ReadFilter.cpp:113-115)
    const uint64_t s = mix64(c.read_seed ^ mix64(rid + 0x632BE59BD9B4E019ULL));
    uint64_t pos = mix64(s + 0x9E3779B97F4A7C15ULL) % c.genome_len;
    const bool rc = mix64(s + 2 * 0x9E3779B97F4A7C15ULL) < c.t_rc;
    uint64_t t = 3, emitted = 0;
    while (emitted < len) {
        const uint64_t r = mix64(s + t * 0x9E3779B97F4A7C15ULL);
        ++t;
        const uint32_t aux = (uint32_t)(mix64(r) >> 40);
        uint32_t code;
        if (r < c.t_ins) {
            code = aux & 3u;                                    // random inserted base
        } else if (r < c.t_del) {
            pos = pos + 1 == c.genome_len ? 0 : pos + 1;        // genome base skipped
            continue;
        } else if (r < c.t_sub) {
            code = (genome_code(c, pos) + 1 + aux % 3u) & 3u;   // one of the three other bases
            pos = pos + 1 == c.genome_len ? 0 : pos + 1;
        } else {
            code = genome_code(c, pos);
            pos = pos + 1 == c.genome_len ? 0 : pos + 1;
        }
        if (rc) out[len - 1 - emitted] = letters[code ^ 1u];    // A<->T, C<->G
        else out[emitted] = letters[code];
        ++emitted;
    }
}

__global__ void __launch_bounds__(128)
synth_kernel(SynthConst c, uint64_t first_read, uint32_t num_reads, const uint64_t *__restrict__ off,
             char *__restrict__ bases) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < num_reads; i += gridDim.x * blockDim.x)
        synth_one_read(c, first_read + i, off[i + 1] - off[i], bases + off[i]);
}

static int make_const(const nsmh_synth_params *p, SynthConst &c) {
    if (!p || p->genome_len == 0) return fail(NSMH_EINVAL, "synth: genome_len must be > 0");
    double pi = p->p_ins, pd = p->p_del, ps = p->p_sub, pr = p->p_rc;
    if (pi < 0 || pd < 0 || ps < 0 || pi + pd + ps >= 1.0 || pd + pi + ps < 0 || pr < 0 || pr > 1)
        return fail(NSMH_EINVAL, "synth: bad probabilities");
    const double two64 = 18446744073709551616.0;
    c.genome_len = p->genome_len;
    c.genome_seed = p->genome_seed;
    c.read_seed = p->read_seed;
    c.t_ins = (uint64_t)(pi * two64);
    c.t_del = (uint64_t)((pi + pd) * two64);
    c.t_sub = (uint64_t)((pi + pd + ps) * two64);
    c.t_rc = pr >= 1.0 ? ~0ULL : (uint64_t)(pr * two64);
    return NSMH_OK;
}

} // namespace nsmh

using namespace nsmh;

extern "C" {

int nsmh_synth_reads_host(const nsmh_synth_params *p, uint64_t first_read, uint32_t num_reads,
                          const uint64_t *offsets, char *bases) {
    SynthConst c;
    NSMH_TRY(make_const(p, c));
    if (num_reads && (!offsets || (!bases && offsets[num_reads] > offsets[0])))
        return fail(NSMH_EINVAL, "synth: null buffers");
    for (uint32_t i = 0; i < num_reads; ++i)
        synth_one_read(c, first_read + i, offsets[i + 1] - offsets[i], bases + offsets[i]);
    return NSMH_OK;
}

int nsmh_synth_reads_device(int device, const nsmh_synth_params *p, uint64_t first_read,
                            uint32_t num_reads, const uint64_t *d_offsets, char *d_bases) {
    SynthConst c;
    NSMH_TRY(make_const(p, c));
    if (num_reads == 0) return NSMH_OK;
    if (!d_offsets || !d_bases) return fail(NSMH_EINVAL, "synth: null buffers");
    int prev = 0;
    NSMH_CK(cudaGetDevice(&prev));
    NSMH_CK(cudaSetDevice(device));
    synth_kernel<<<(num_reads + 127) / 128, 128>>>(c, first_read, num_reads, d_offsets, d_bases);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, "synth_kernel", __FILE__, __LINE__);
    return NSMH_OK;
}

} // extern "C"
