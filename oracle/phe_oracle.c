/*
 * phe_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the checker for the B200 hand-evaluation / equity path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under texasholdempocker-rl_b200/ links, imports or
 * calls it, and none of its tables are handed to the CUDA library.
 *
 * PARITY STATUS: the reference (weipeilun/texasholdempocker-rl) keeps the
 * arithmetic of this path in the un-vendored, un-pinned PyPI dependency
 * `phevaluator` (HenryRLee/PokerHandEvaluator, python/ sub-package; call site
 * env/game.py:10,675) and its tree holds no golden vector for ranks, winners or
 * equities ("parity unpinned" by the reference itself).  The
 * oracle is therefore pinned three other ways (tests/test_oracle_*.py):
 *   1. upstream README examples 292 / 236 and the canonical 7462-class
 *      numbering with category bounds 10/166/322/1599/1609/2467/3325/6185/7462;
 *   2. the published 7-card category histogram over all C(52,7) hands;
 *   3. outputs of the reference's own code run in the build container:
 *      GameEnv.get_winner_set (legacy in-tree evaluator, env/game.py:635-659,
 *      valid for non-wheel hands), GameEnv.get_winner_set_v2 /
 *      generate_game_result / simulate_processes / new_random / finish_game
 *      driven through a phevaluator shim (tests/golden/make_golden.py).
 *
 * Two evaluators live here and must agree on every input:
 *   A. orc_rank7_brute  -- first principles: best of the 21 five-card subsets,
 *      each looked up in a class list built by sorting the strength codes of all
 *      C(52,5) hands (no formulas for class numbers; 7462 falls out).
 *   B. orc_rank7        -- restatement of phevaluator's published 7-card
 *      algorithm (env/game.py:675 -> phevaluator.evaluate_cards): suit hash ->
 *      SUITS[]; flush -> FLUSH[8192]; otherwise 13-digit base-5 rank-count
 *      vector -> hash_quinary via DP[5][14][10] -> NO_FLUSH_7[49205].  Its table
 *      payloads are filled from evaluator A.
 *
 * Card ids follow utils/pokerhandevaluator_utils.py:3-4: id = figure*4 + decor,
 * figure 0 ('2') .. 12 ('A'), decor 0 C, 1 D, 2 H, 3 S (env/constants.py:14-36).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define NCLASS 7462

/* ------------------------------------------------------------------ */
/* A. first-principles evaluator                                        */
/* ------------------------------------------------------------------ */

static uint32_t g_codes[NCLASS + 8]; /* strength codes, strongest first */
static int g_ncodes = 0;

/* strength code of a 5-card hand: category in bits 20.., tie-break nibbles below */
static uint32_t strength5(const uint8_t c[5])
{
    int cnt[13] = {0};
    int flush = 1;
    for (int i = 0; i < 5; i++) {
        cnt[c[i] >> 2]++;
        if ((c[i] & 3) != (c[0] & 3)) flush = 0;
    }
    /* distinct ranks ordered by (multiplicity desc, rank desc) */
    int seq[5], nseq = 0;
    for (int m = 4; m >= 1; m--)
        for (int r = 12; r >= 0; r--)
            if (cnt[r] == m) seq[nseq++] = r;
    int shape0 = cnt[seq[0]];
    int shape1 = nseq > 1 ? cnt[seq[1]] : 0;
    int straight = 0, high = 0;
    if (nseq == 5) {
        if (seq[0] - seq[4] == 4) { straight = 1; high = seq[0]; }
        else if (seq[0] == 12 && seq[1] == 3 && seq[2] == 2 && seq[3] == 1 && seq[4] == 0) {
            straight = 1; high = 3; /* wheel: A-2-3-4-5 is five-high */
        }
    }
    int cat;
    if (straight && flush) cat = 8;
    else if (shape0 == 4) cat = 7;
    else if (shape0 == 3 && shape1 == 2) cat = 6;
    else if (flush) cat = 5;
    else if (straight) cat = 4;
    else if (shape0 == 3) cat = 3;
    else if (shape0 == 2 && shape1 == 2) cat = 2;
    else if (shape0 == 2) cat = 1;
    else cat = 0;
    uint32_t code = (uint32_t)cat << 20;
    if (straight) {
        code |= (uint32_t)high << 16;
    } else {
        for (int i = 0; i < nseq; i++) code |= (uint32_t)seq[i] << (16 - 4 * i);
    }
    return code;
}

static int cmp_u32_desc(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}

static void build_classes(void)
{
    /* all C(52,5) = 2,598,960 hands -> codes -> sort -> unique */
    size_t cap = 2598960;
    uint32_t *all = (uint32_t *)malloc(cap * sizeof(uint32_t));
    size_t n = 0;
    uint8_t c[5];
    for (c[0] = 0; c[0] < 48; c[0]++)
        for (c[1] = c[0] + 1; c[1] < 49; c[1]++)
            for (c[2] = c[1] + 1; c[2] < 50; c[2]++)
                for (c[3] = c[2] + 1; c[3] < 51; c[3]++)
                    for (c[4] = c[3] + 1; c[4] < 52; c[4]++)
                        all[n++] = strength5(c);
    qsort(all, n, sizeof(uint32_t), cmp_u32_desc);
    g_ncodes = 0;
    for (size_t i = 0; i < n; i++)
        if (i == 0 || all[i] != all[i - 1]) {
            if (g_ncodes < NCLASS + 8) g_codes[g_ncodes] = all[i];
            g_ncodes++;
        }
    free(all);
}

static int class_of_code(uint32_t code)
{
    int lo = 0, hi = g_ncodes - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (g_codes[mid] == code) return mid + 1;
        if (g_codes[mid] > code) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

int orc_num_classes(void) { return g_ncodes; }

int orc_rank5_brute(const uint8_t c[5]) { return class_of_code(strength5(c)); }

int orc_rank7_brute(const uint8_t c[7])
{
    int best = 1 << 30;
    uint8_t h[5];
    for (int a = 0; a < 7; a++)
        for (int b = a + 1; b < 7; b++) {
            int k = 0;
            for (int i = 0; i < 7; i++)
                if (i != a && i != b) h[k++] = c[i];
            int r = class_of_code(strength5(h));
            if (r < best) best = r;
        }
    return best;
}

/* ------------------------------------------------------------------ */
/* B. restated phevaluator 7-card algorithm                             */
/* ------------------------------------------------------------------ */

static uint8_t  T_SUITS[4609];
static uint16_t T_FLUSH[8192];
static uint32_t T_DP[5][14][10];
static uint16_t T_NOFLUSH7[49205];
static const uint16_t SUITBIT[4] = {0x1, 0x8, 0x40, 0x200};

static uint32_t hash_quinary(const uint8_t q[13], int k)
{
    uint32_t sum = 0;
    for (int i = 0; i < 13; i++) {
        if (q[i]) {
            sum += T_DP[q[i]][13 - i - 1][k];
            k -= q[i];
            if (k <= 0) break;
        }
    }
    return sum;
}

static void fill_noflush(uint8_t q[13], int pos, int left)
{
    if (pos == 13) {
        if (left) return;
        uint8_t c[7]; int n = 0;
        /* round-robin suits: same-rank cards get distinct suits, no suit holds > 2 cards */
        for (int r = 0; r < 13; r++)
            for (int j = 0; j < q[r]; j++) { c[n] = (uint8_t)(r * 4 + (n & 3)); n++; }
        T_NOFLUSH7[hash_quinary(q, 7)] = (uint16_t)orc_rank7_brute(c);
        return;
    }
    for (int d = 0; d <= 4 && d <= left; d++) { q[pos] = (uint8_t)d; fill_noflush(q, pos + 1, left - d); }
    q[pos] = 0;
}

static void build_tables(void)
{
    /* SUITS: index = sum of SUITBIT over cards (octal digit per suit) */
    memset(T_SUITS, 0, sizeof T_SUITS);
    for (int a = 0; a <= 7; a++)
        for (int b = 0; a + b <= 7; b++)
            for (int c = 0; a + b + c <= 7; c++)
                for (int d = 0; a + b + c + d <= 7; d++) {
                    int idx = a * 1 + b * 8 + c * 64 + d * 512;
                    int s = 0;
                    if (a >= 5) s = 1; else if (b >= 5) s = 2; else if (c >= 5) s = 3; else if (d >= 5) s = 4;
                    T_SUITS[idx] = (uint8_t)s;
                }
    /* DP[l][n][k] = #(d0 < l, then n digits in 0..4) with d0 + digits = k */
    static uint32_t f[15][10]; /* f[n][k] = # n-digit strings (digits 0..4) with sum k */
    memset(f, 0, sizeof f);
    f[0][0] = 1;
    for (int n = 1; n <= 14; n++)
        for (int k = 0; k < 10; k++)
            for (int d = 0; d <= 4 && d <= k; d++) f[n][k] += f[n - 1][k - d];
    memset(T_DP, 0, sizeof T_DP);
    for (int l = 1; l < 5; l++)
        for (int n = 0; n < 14; n++)
            for (int k = 0; k < 10; k++)
                for (int d0 = 0; d0 < l && d0 <= k; d0++) T_DP[l][n][k] += f[n][k - d0];
    /* FLUSH[mask]: 5..7 suited ranks, padded with off-suit deuces/treys that cannot matter */
    memset(T_FLUSH, 0, sizeof T_FLUSH);
    for (int m = 0; m < 8192; m++) {
        int pc = __builtin_popcount(m);
        if (pc < 5 || pc > 7) continue;
        uint8_t c[7]; int n = 0;
        for (int r = 0; r < 13; r++) if (m >> r & 1) c[n++] = (uint8_t)(r * 4 + 3);
        /* fill with cards of other suits whose ranks are not in m where possible */
        for (int r = 0; r < 13 && n < 7; r++)
            if (!(m >> r & 1)) { c[n] = (uint8_t)(r * 4 + (n & 1)); n++; }
        T_FLUSH[m] = (uint16_t)orc_rank7_brute(c);
    }
    /* NO_FLUSH_7: every 13-digit count vector with sum 7, digits <= 4 */
    memset(T_NOFLUSH7, 0, sizeof T_NOFLUSH7);
    uint8_t q[13];
    memset(q, 0, sizeof q);
    fill_noflush(q, 0, 7);
}

int orc_rank7(const uint8_t c[7])
{
    int suit_hash = 0;
    for (int i = 0; i < 7; i++) suit_hash += SUITBIT[c[i] & 3];
    int fs = T_SUITS[suit_hash];
    if (fs) {
        int sb[4] = {0, 0, 0, 0};
        for (int i = 0; i < 7; i++) sb[c[i] & 3] |= 1 << (c[i] >> 2);
        return T_FLUSH[sb[fs - 1]];
    }
    uint8_t q[13];
    memset(q, 0, sizeof q);
    for (int i = 0; i < 7; i++) q[c[i] >> 2]++;
    return T_NOFLUSH7[hash_quinary(q, 7)];
}

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void do_init(void) { build_classes(); build_tables(); }
int orc_init(void) { pthread_once(&g_once, do_init); return g_ncodes == NCLASS ? 0 : -1; }

/* copy out the restated tables (for the pure-Python phevaluator shim) */
void orc_get_tables(uint8_t *suits4609, uint16_t *flush8192, uint32_t *dp_5_14_10, uint16_t *noflush49205)
{
    orc_init();
    memcpy(suits4609, T_SUITS, sizeof T_SUITS);
    memcpy(flush8192, T_FLUSH, sizeof T_FLUSH);
    memcpy(dp_5_14_10, T_DP, sizeof T_DP);
    memcpy(noflush49205, T_NOFLUSH7, sizeof T_NOFLUSH7);
}

void orc_rank7_batch(const uint8_t *cards, int64_t n, uint16_t *out, int brute)
{
    orc_init();
    for (int64_t i = 0; i < n; i++)
        out[i] = (uint16_t)(brute ? orc_rank7_brute(cards + 7 * i) : orc_rank7(cards + 7 * i));
}

/* mask layout of the B200 library: bit 16*suit + rank */
void orc_rank7_masks(const uint64_t *masks, int64_t n, uint16_t *out)
{
    orc_init();
    for (int64_t i = 0; i < n; i++) {
        uint8_t c[7]; int k = 0;
        uint64_t m = masks[i];
        while (m && k < 7) {
            int b = __builtin_ctzll(m); m &= m - 1;
            c[k++] = (uint8_t)((b & 15) * 4 + (b >> 4));
        }
        out[i] = (uint16_t)orc_rank7(c);
    }
}

static int category_of_rank(int r)
{
    /* K2 bounds */
    static const int bound[8] = {10, 166, 322, 1599, 1609, 2467, 3325, 6185};
    for (int i = 0; i < 8; i++)
        if (r <= bound[i]) return i;
    return 8;
}

/* ------------------------------------------------------------------ */
/* exhaustive C(52,7): colex combination index range [begin,end)        */
/* ------------------------------------------------------------------ */

static uint64_t binom(int n, int k)
{
    if (k < 0 || n < k) return 0;
    uint64_t r = 1;
    for (int i = 1; i <= k; i++) r = r * (uint64_t)(n - k + i) / (uint64_t)i;
    return r;
}

/* colex unrank: idx = sum_i C(c_i, i+1), c_0 < c_1 < ... < c_6 */
static void colex_unrank7(uint64_t idx, uint8_t c[7])
{
    for (int i = 6; i >= 0; i--) {
        int v = i;
        while (binom(v + 1, i + 1) <= idx) v++;
        c[i] = (uint8_t)v;
        idx -= binom(v, i + 1);
    }
}

/* advance to the colex successor */
static void colex_next7(uint8_t c[7])
{
    int i = 0;
    while (i < 6 && c[i] + 1 == c[i + 1]) { c[i] = (uint8_t)i; i++; }
    c[i]++;
}

typedef struct { uint64_t begin, end; uint64_t h9[9]; uint64_t *h7463; } enum_job;

static void *enum_worker(void *p)
{
    enum_job *j = (enum_job *)p;
    if (j->begin >= j->end) return 0;
    uint8_t c[7];
    colex_unrank7(j->begin, c);
    for (uint64_t i = j->begin; i < j->end; i++) {
        int r = orc_rank7(c);
        j->h9[category_of_rank(r)]++;
        j->h7463[r]++;
        colex_next7(c);
    }
    return 0;
}

void orc_enum7(uint64_t begin, uint64_t end, uint64_t *hist9, uint64_t *hist7463, int nthreads)
{
    orc_init();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    enum_job *jobs = (enum_job *)calloc((size_t)nthreads, sizeof(enum_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    uint64_t total = end > begin ? end - begin : 0;
    for (int t = 0; t < nthreads; t++) {
        jobs[t].begin = begin + total * (uint64_t)t / (uint64_t)nthreads;
        jobs[t].end = begin + total * (uint64_t)(t + 1) / (uint64_t)nthreads;
        jobs[t].h7463 = (uint64_t *)calloc(NCLASS + 1, sizeof(uint64_t));
        pthread_create(&th[t], 0, enum_worker, &jobs[t]);
    }
    memset(hist9, 0, 9 * sizeof(uint64_t));
    memset(hist7463, 0, (NCLASS + 1) * sizeof(uint64_t));
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], 0);
        for (int i = 0; i < 9; i++) hist9[i] += jobs[t].h9[i];
        for (int i = 0; i <= NCLASS; i++) hist7463[i] += jobs[t].h7463[i];
        free(jobs[t].h7463);
    }
    free(jobs); free(th);
}

void orc_colex_unrank7(uint64_t idx, uint8_t *c) { colex_unrank7(idx, c); }

/* ------------------------------------------------------------------ */
/* showdown: env/game.py:663-681 (get_winner_set_v2)                    */
/* ------------------------------------------------------------------ */

/* holes: P x 2 card ids; live: bit p set if player p is in the showdown dict.
 * Returns the bitmask of winners (all live players of minimal rank).  A single
 * live player wins without evaluation (env/game.py:664-665). */
uint32_t orc_winner_mask(const uint8_t board[5], const uint8_t *holes, int P, uint32_t live, uint16_t *ranks_out)
{
    orc_init();
    if (__builtin_popcount(live) == 1) {
        if (ranks_out) for (int p = 0; p < P; p++) ranks_out[p] = 0;
        return live;
    }
    int lowest = 1 << 30; uint32_t win = 0;
    for (int p = 0; p < P; p++) {
        if (ranks_out) ranks_out[p] = 0;
        if (!(live >> p & 1)) continue;
        uint8_t c[7];
        memcpy(c, board, 5); c[5] = holes[2 * p]; c[6] = holes[2 * p + 1];
        int r = orc_rank7(c);
        if (ranks_out) ranks_out[p] = (uint16_t)r;
        if (r <= lowest) {
            if (r < lowest) { lowest = r; win = 0; }
            win |= 1u << p;
        }
    }
    return win;
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) -- the shared seed stream       */
/* ------------------------------------------------------------------ */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The dealing stream the CUDA kernels use (DESIGN.md "RNG stream"), restated.
 *   A deal is a sequence of slots: at most one single card, then pairs.  Philox block b = philox4x32_10(counter =
 *   (sample lo32, sample hi32, stream id, b), key = (seed lo32, seed hi32)); its four words give eight 16-bit draws,
 *   low half first.
 *   single slot (if asked for): draws 0 and 1 of block 0 are reserved for it; draw r is valid if r < 1260*52 and
 *     proposes card id r % 52; the first valid proposal whose card is free is taken; if neither works, further
 *     draws come from blocks 0x80000000, 0x80000001, ...
 *   pair slots, in order: draws continue after the reserved word (from draw 0 if there is no single); r is valid if
 *     r < 49*1326 and proposes the pair with index r % 1326 (index C(c1,2)+c0 over ids c0 < c1), accepted iff both
 *     cards are free.  Accepted slots mark their cards used.
 *   Hold'em roles: single + first pairs = missing board cards, remaining pairs = opponents in order. */
static inline uint8_t bit_to_id(int b) { return (uint8_t)((b & 15) * 4 + (b >> 4)); }
static inline int id_to_bit(int id) { return 16 * (id & 3) + (id >> 2); }

static int mask_to_ids(uint64_t m, uint8_t *out, int cap)
{
    int k = 0;
    while (m && k < cap) { int b = __builtin_ctzll(m); m &= m - 1; out[k++] = bit_to_id(b); }
    return k;
}

static void pair_from_index(uint32_t q, int *c0, int *c1)
{
    int hi = 1;
    while ((uint32_t)((hi + 1) * hi / 2) <= q) hi++;
    *c1 = hi; *c0 = (int)q - hi * (hi - 1) / 2;
}

/* deals `want_single` + n_pairs slots; ids_out = [single?][pair0 lo, pair0 hi][pair1 ...] (pairs as sorted ids) */
static void deal_shared(uint64_t known, uint64_t seed, uint64_t sample, uint32_t stream, int want_single, int n_pairs, uint8_t *ids_out)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)sample, (uint32_t)(sample >> 32), stream, 0};
    uint32_t out[4];
    uint64_t used = known;
    int n = 0;
    orc_philox4x32_10(ctr, key, out);
    if (want_single) {
        int have = 0;
        for (int h = 0; h < 2 && !have; h++) {
            uint32_t r = (out[0] >> (16 * h)) & 0xFFFFu;
            int c = (int)(r % 52u);
            if (r < 1260u * 52u && !(used >> id_to_bit(c) & 1)) { have = 1; used |= 1ull << id_to_bit(c); ids_out[n++] = (uint8_t)c; }
        }
        for (uint32_t t = 0; !have; t++) {
            uint32_t c2[4] = {ctr[0], ctr[1], ctr[2], 0x80000000u + t}, ex[4];
            orc_philox4x32_10(c2, key, ex);
            for (int d = 0; d < 8 && !have; d++) {
                uint32_t r = (ex[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
                int c = (int)(r % 52u);
                if (r < 1260u * 52u && !(used >> id_to_bit(c) & 1)) { have = 1; used |= 1ull << id_to_bit(c); ids_out[n++] = (uint8_t)c; }
            }
        }
    }
    int got = 0, first_word = want_single ? 1 : 0;
    for (uint32_t blk = 0; got < n_pairs; blk++) {
        if (blk) { ctr[3] = blk; orc_philox4x32_10(ctr, key, out); }
        for (int wi = first_word; wi < 4 && got < n_pairs; wi++)
            for (int h = 0; h < 2 && got < n_pairs; h++) {
                uint32_t r = (out[wi] >> (16 * h)) & 0xFFFFu;
                if (r >= 49u * 1326u) continue;
                int c0, c1;
                pair_from_index(r % 1326u, &c0, &c1);
                uint64_t m = (1ull << id_to_bit(c0)) | (1ull << id_to_bit(c1));
                if (m & used) continue;
                used |= m;
                ids_out[n++] = (uint8_t)c0; ids_out[n++] = (uint8_t)c1;
                got++;
            }
        first_word = 0;
    }
}

void orc_deal(uint64_t known, uint64_t seed, uint64_t sample, uint32_t stream, int need, uint8_t *ids_out)
{
    deal_shared(known, seed, sample, stream, need & 1, need >> 1, ids_out);
}

/* Monte-Carlo equity of one state with the shared stream; bit-exact partner of
 * phe_equity_mc.  Semantics per rollout follow env/workflow.py:130-148 with P-1
 * opponents: hero WIN if sole minimum rank, EVEN (tie) if in a winner set of
 * size >= 2 (env/game.py:551-558), else LOSE. */
void orc_equity_mc(uint64_t known, uint64_t hero, int n_opp, uint64_t rollouts, uint64_t sample_offset,
                   uint64_t seed, uint32_t state_idx, uint64_t wtl[3])
{
    orc_init();
    uint8_t hero_ids[2], board_ids[5];
    mask_to_ids(hero, hero_ids, 2);
    int nb = mask_to_ids(known & ~hero, board_ids, 5);
    int nbm = 5 - nb;
    wtl[0] = wtl[1] = wtl[2] = 0;
    for (uint64_t s = 0; s < rollouts; s++) {
        uint8_t dealt[32];
        deal_shared(known, seed, sample_offset + s, state_idx, nbm & 1, nbm / 2 + n_opp, dealt);
        uint8_t c[7];
        int k = 0;
        for (int i = 0; i < nb; i++) c[i] = board_ids[i];
        for (int i = nb; i < 5; i++) c[i] = dealt[k++];      /* single + board pairs complete the board */
        c[5] = hero_ids[0]; c[6] = hero_ids[1];
        int hr = orc_rank7(c);
        int best = 1 << 30;
        for (int o = 0; o < n_opp; o++) {
            c[5] = dealt[k++]; c[6] = dealt[k++];
            int r = orc_rank7(c);
            if (r < best) best = r;
        }
        if (hr < best) wtl[0]++; else if (hr == best) wtl[1]++; else wtl[2]++;
    }
}

/* statistically independent Monte Carlo (xorshift128+ and a partial
 * Fisher-Yates over card ids) for the 3-sigma cross-check */
static uint64_t xs_next(uint64_t s[2])
{
    uint64_t a = s[0], b = s[1];
    s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}

void orc_equity_mc_indep(uint64_t known, uint64_t hero, int n_opp, uint64_t rollouts, uint64_t seed, uint64_t wtl[3])
{
    orc_init();
    uint8_t hero_ids[2], board_ids[5], deck[52];
    mask_to_ids(hero, hero_ids, 2);
    int nb = mask_to_ids(known & ~hero, board_ids, 5);
    int nd = 0;
    for (int id = 0; id < 52; id++) if (!(known >> id_to_bit(id) & 1)) deck[nd++] = (uint8_t)id;
    int need = (5 - nb) + 2 * n_opp;
    uint64_t st[2] = {seed * 0x9E3779B97F4A7C15ull + 1, seed ^ 0xD1B54A32D192ED03ull};
    for (int i = 0; i < 16; i++) xs_next(st);
    wtl[0] = wtl[1] = wtl[2] = 0;
    for (uint64_t s = 0; s < rollouts; s++) {
        for (int i = 0; i < need; i++) {
            /* unbiased bounded draw by rejection on the top bits */
            uint64_t range = (uint64_t)(nd - i), lim = UINT64_MAX - UINT64_MAX % range, r;
            do { r = xs_next(st); } while (r >= lim);
            int j = i + (int)(r % range);
            uint8_t t = deck[i]; deck[i] = deck[j]; deck[j] = t;
        }
        uint8_t c[7]; int k = 0;
        for (int i = 0; i < nb; i++) c[i] = board_ids[i];
        for (int i = nb; i < 5; i++) c[i] = deck[k++];
        c[5] = hero_ids[0]; c[6] = hero_ids[1];
        int hr = orc_rank7(c), best = 1 << 30;
        for (int o = 0; o < n_opp; o++) {
            c[5] = deck[k++]; c[6] = deck[k++];
            int r = orc_rank7(c);
            if (r < best) best = r;
        }
        if (hr < best) wtl[0]++; else if (hr == best) wtl[1]++; else wtl[2]++;
    }
}

/* ------------------------------------------------------------------ */
/* exact-plan equity: env/workflow.py:150-199                           */
/* ------------------------------------------------------------------ */

/* plan steps (gen_num, num_random, is_segment_end) as in
 * config/default.yml:73-133.  exist = hero(2) + known board in deal order
 * (env/workflow.py:212-221).  Enumeration visits deck indices (= card ids,
 * env/game.py:14-18) in increasing order, continuing after the previous card
 * inside a segment and restarting at 0 after a segment end (:174-177).  A step
 * with gen_num > 1 samples num_random completions and must be last (:181-195).
 * The random step uses the shared stream with sample index = the running leaf
 * number of the enclosing enumeration * num_random + draw number. */
typedef struct {
    const int32_t *plan; int nsteps;
    uint64_t seed; uint32_t stream;
    uint64_t w, t, l, n, rnd_leaf;
} plan_ctx;

static void plan_leaf(plan_ctx *cx, const uint8_t *cards /* 9 ids, layout workflow.py:131-135 */)
{
    uint8_t c[7];
    c[0] = cards[2]; c[1] = cards[3]; c[2] = cards[4]; c[3] = cards[5]; c[4] = cards[6];
    c[5] = cards[0]; c[6] = cards[1];
    int hr = orc_rank7(c);
    c[5] = cards[7]; c[6] = cards[8];
    int vr = orc_rank7(c);
    if (hr < vr) cx->w++; else if (hr == vr) cx->t++; else cx->l++;
    cx->n++;
}

static void plan_rec(plan_ctx *cx, int step, uint8_t *cards, int ncards, int start_idx)
{
    int gen = cx->plan[3 * step], nrand = cx->plan[3 * step + 1], seg_end = cx->plan[3 * step + 2];
    int last = (step == cx->nsteps - 1);
    uint64_t exist = 0;
    for (int i = 0; i < ncards; i++) exist |= 1ull << cards[i];
    if (gen == 1) {
        int more_in_segment = 0;
        if (!seg_end)
            for (int s = step + 1; s < cx->nsteps; s++) {
                more_in_segment += cx->plan[3 * s];
                if (cx->plan[3 * s + 2]) break;
            }
        for (int idx = start_idx; idx < 52; idx++) {
            if (exist >> idx & 1) continue;
            cards[ncards] = (uint8_t)idx;
            if (last) plan_leaf(cx, cards);
            else if (more_in_segment + idx < 52)
                plan_rec(cx, step + 1, cards, ncards + 1, seg_end ? 0 : idx + 1);
        }
    } else {
        uint64_t known_bits = 0;
        for (int i = 0; i < ncards; i++) known_bits |= 1ull << id_to_bit(cards[i]);
        for (int j = 0; j < nrand; j++) {
            /* [single?][pairs...]: the last pair is the villain's hand, everything before it completes the board */
            deal_shared(known_bits, cx->seed, cx->rnd_leaf * (uint64_t)nrand + (uint64_t)j, cx->stream, gen & 1, gen >> 1, cards + ncards);
            plan_leaf(cx, cards);
        }
        cx->rnd_leaf++;
    }
}

/* returns 0 on success; out = {wins, ties, losses, n} ; game_result_sum = wins - losses */
int orc_equity_plan(const uint8_t *exist_ids, int n_exist, const int32_t *plan, int nsteps,
                    uint64_t seed, uint32_t stream, uint64_t out[4])
{
    orc_init();
    int total = n_exist;
    for (int s = 0; s < nsteps; s++) total += plan[3 * s];
    if (total != 9) return -1;
    for (int s = 0; s < nsteps - 1; s++) if (plan[3 * s] != 1) return -2;
    plan_ctx cx; memset(&cx, 0, sizeof cx);
    cx.plan = plan; cx.nsteps = nsteps; cx.seed = seed; cx.stream = stream;
    uint8_t cards[16];
    memcpy(cards, exist_ids, (size_t)n_exist);
    plan_rec(&cx, 0, cards, n_exist, 0);
    out[0] = cx.w; out[1] = cx.t; out[2] = cx.l; out[3] = cx.n;
    return 0;
}

/* ------------------------------------------------------------------ */
/* throughput probes for bench.py's cpu_baseline                        */
/* ------------------------------------------------------------------ */

typedef struct { const uint8_t *cards; int64_t n; uint64_t acc; } bench_job;
static void *bench_worker(void *p)
{
    bench_job *j = (bench_job *)p; uint64_t a = 0;
    for (int64_t i = 0; i < j->n; i++) a += (uint64_t)orc_rank7(j->cards + 7 * i);
    j->acc = a; return 0;
}

/* evaluate n hands with `nthreads` threads; returns a checksum (sum of ranks) */
uint64_t orc_rank7_sum_mt(const uint8_t *cards, int64_t n, int nthreads)
{
    orc_init();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    bench_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < nthreads; t++) {
        int64_t b = n * t / nthreads, e = n * (t + 1) / nthreads;
        jobs[t].cards = cards + 7 * b; jobs[t].n = e - b; jobs[t].acc = 0;
        pthread_create(&th[t], 0, bench_worker, &jobs[t]);
    }
    uint64_t a = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], 0); a += jobs[t].acc; }
    return a;
}
