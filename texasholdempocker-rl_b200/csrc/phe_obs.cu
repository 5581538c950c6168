// phe_obs.cu -- batched observation encoder: the int32[93] feature vector of Env.get_obs
// (env/env.py:416-581: get_obs, get_all_historical_action, get_all_player_current_status, _map_obs_idx_to_model_index;
// cutters env/env.py:52-81 over tools/biner.py CutByThreshold; card slots env/env.py:702-723).
//
// One thread per observation.  All ratio features are an IEEE double division of two chip counts compared against the
// reference's own threshold lists (passed in by the host layer exactly as numpy computed them), so the bins are
// bit-identical to the reference's Python floats.  Integer / byte work: 131 + 8*n_bins bytes read and 4*F bytes
// written per observation.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/phe_b200.h"

void phe_internal_count_launch();          // phe_kernels.cu

namespace {

struct CutTab {
    double thr[PHE_OBS_CUTTERS][PHE_OBS_MAX_THR];
    int len[PHE_OBS_CUTTERS];
};

// tools/biner.py:8-27.  `invalid`: the cutter was built with include_invalid_bin and marker < 0 selects bin len + 1.
__device__ __forceinline__ int cut(const CutTab& t, int c, long long num, long long den, bool with_marker, long long marker)
{
    const int n = t.len[c];
    if (with_marker && marker < 0) return n + 1;
    if (den == 0) return n;
    const double v = (double)num / (double)den;
    int b = n;
    for (int i = n - 1; i >= 0; i--)
        if (v < t.thr[c][i]) b = i;             // first threshold above the value (lists are increasing, scan keeps the lowest)
    return b;
}

__global__ void __launch_bounds__(128)
k_encode_obs(const int32_t* __restrict__ round_role, const uint8_t* __restrict__ cards, const long long* __restrict__ values,
             const long long* __restrict__ bins, const long long* __restrict__ others, const uint8_t* __restrict__ history,
             const CutTab* __restrict__ g_tab, int64_t n, int n_bins, int n_other, int n_hist, int32_t* __restrict__ out)
{
    __shared__ CutTab tab;
    for (int i = threadIdx.x; i < (int)(sizeof(CutTab) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&tab)[i] = reinterpret_cast<const uint32_t*>(g_tab)[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int F = 2 + 14 + (10 + 4 * n_bins) + 3 * n_other + n_hist;
    int32_t* o = out + i * F;
    int k = 0;
    o[k++] = round_role[2 * i];
    o[k++] = round_role[2 * i + 1];
    // card slots: id 0..51 -> (figure, decor); PHE_OBS_CARD_NONE -> (14, 5); PHE_OBS_CARD_HIDDEN -> (13, 4)
    for (int c = 0; c < 7; c++) {
        const uint32_t id = cards[7 * i + c];
        o[k + c] = id < 52 ? (int)(id >> 2) : (id == PHE_OBS_CARD_HIDDEN ? 13 : 14);
        o[k + 7 + c] = id < 52 ? (int)(id & 3) : (id == PHE_OBS_CARD_HIDDEN ? 4 : 5);
    }
    k += 14;
    const long long game_init = values[5 * i], pot = values[5 * i + 1], a_init = values[5 * i + 2], a_left = values[5 * i + 3],
                    a_bet = values[5 * i + 4];
    const long long asset[3] = {game_init, a_init, a_left};
    // env/env.py:539-547
    o[k++] = cut(tab, 0, a_init, game_init, false, 0);
    o[k++] = cut(tab, 1, a_left, game_init, false, 0);
    o[k++] = cut(tab, 2, a_left, a_init, false, 0);
    o[k++] = cut(tab, 3, a_bet, pot, false, 0);
    for (int a = 0; a < 3; a++) o[k++] = cut(tab, 4 + a, a_bet, asset[a], false, 0);
    for (int a = 0; a < 3; a++) o[k++] = cut(tab, 7 + a, pot, asset[a], false, 0);
    // env/env.py:549-553: the action value of every bin against the pot and the three assets
    for (int b = 0; b < n_bins; b++) {
        const long long v = bins[(int64_t)n_bins * i + b];
        o[k + b] = cut(tab, 10, v, pot, true, v);
    }
    k += n_bins;
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < n_bins; b++) {
            const long long v = bins[(int64_t)n_bins * i + b];
            o[k + b] = cut(tab, 11 + a, v, asset[a], true, v);
        }
        k += n_bins;
    }
    // env/env.py:566-581 + :399-401: per other player (status, left/init bin, init/acting-init bin), field-major
    for (int f = 0; f < 3; f++) {
        for (int j = 0; j < n_other; j++) {
            const long long* op = others + ((int64_t)n_other * i + j) * 3;      // status, left, init
            int v;
            if (f == 0) v = (int)op[0];
            else if (f == 1) v = cut(tab, 14, op[1], op[2], false, 0);
            else v = cut(tab, 15, op[2], a_init, false, 0);
            o[k++] = v;
        }
    }
    for (int h = 0; h < n_hist; h++) o[k++] = history[(int64_t)n_hist * i + h];
}

}  // namespace

extern "C" {

int phe_obs_num_fields(int n_bins, int n_other, int n_hist) { return 2 + 14 + (10 + 4 * n_bins) + 3 * n_other + n_hist; }

int phe_encode_obs(const int32_t* round_role, const uint8_t* cards, const int64_t* values, const int64_t* bins, const int64_t* others,
                   const uint8_t* history, const void* cut_table, int64_t n, int n_bins, int n_other, int n_hist, int32_t* obs_out,
                   void* stream)
{
    if (n < 0 || n_bins < 0 || n_other < 0 || n_hist < 0) return PHE_EINVAL;
    if (n == 0) return PHE_OK;
    if (!round_role || !cards || !values || !cut_table || !obs_out || (n_bins && !bins) || (n_other && !others) || (n_hist && !history))
        return PHE_EINVAL;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    k_encode_obs<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(round_role, cards, reinterpret_cast<const long long*>(values),
                                                                        reinterpret_cast<const long long*>(bins),
                                                                        reinterpret_cast<const long long*>(others), history,
                                                                        static_cast<const CutTab*>(cut_table), n, n_bins, n_other, n_hist, obs_out);
    phe_internal_count_launch();
    return cudaGetLastError() == cudaSuccess ? PHE_OK : PHE_ECUDA;
}

uint64_t phe_obs_cut_table_bytes(void) { return sizeof(CutTab); }

}  // extern "C"
