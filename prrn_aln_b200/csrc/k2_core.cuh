// k2_core.cuh -- per-lane recurrence, direction bits and traceback of kernel K2 (pairwise alignment
// with path for two single sequences: alignC<DPunit>, reference src/fwd2c.h:359-482,670-677;
// src/fwd2c.cc:52-102; src/vmf.cc:103-119).  Shared by the CUDA kernels (k2_align.cu) and the host
// emulation (tests/host_emul/k2_emul.cc).
//
// Same machine mapping as K1 (k1_core.cuh): lane t owns R rows, walks the columns one step behind
// lane t-1; drifted integer values.  In addition every cell emits 4 direction bits
//     bits 0-1  source of H:  0 diagonal, 1 vertical state G, 2 horizontal state F
//     bit  2    G(m,n) was opened from H(m-1,n)   (else extended from G(m-1,n))
//     bit  3    F(m,n) was opened from H(m,n-1)   (else extended from F(m,n-1))
// with the reference's tie rules (fwd2c.h:405-453): a gap opens only if strictly better than
// extending, F beats G on ties, and a gap state replaces the diagonal only if strictly better.
// One lane-step = 16 cells = one 64-bit word, stored in wavefront order
//     word[(pass * (LS + 31) + step) * 32 + lane],  step = n + lane,  nibble k = row within the lane
// so that a warp's store is one fully coalesced 256-byte line.
//
// The traceback walks the bits back from (LQ-1, LS-1), then replays the path forward applying the
// reference's record rules (update(): fwd2c.cc:93-102; Vmf::add at NEWD/NEWV/NEWH cells,
// fwd2c.h:465-467) so that the corner list equals what Vmf::traceback returns, quirks included
// (e.g. the NEWV record of a vertical run survives only when the run has length 1).
#pragma once
#include "k1_core.cuh"

// TraceBackDir (src/aln.h:47-52)
enum { K2_DIAG = 2, K2_NEWD = 3, K2_VERT = 4, K2_HORI = 8, K2_NEWV = 12, K2_NEWH = 13 };
PG_HD bool k2_isdiag(int d) { return d == K2_DIAG || d == K2_NEWD; }
PG_HD bool k2_isvert(int d) { return d == K2_VERT || d == K2_NEWV; }
PG_HD bool k2_ishori(int d) { return d == K2_HORI || d == K2_NEWH; }

template <int R>
struct K2Lane {
    int H[R];       // Ht(mbase+k, n-1)
    int E[R];       // horizontal state for the coming column (eager)
    unsigned long long eopen;   // bit 4k + 3 (its place in the direction word): that state was opened (not extended) when it was computed
    int hdiag;      // Ht(mbase-1, n-1)
};

template <int R>
PG_HD void k2_lane_init(K2Lane<R>& L, const K1Geom& g, int mbase)
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        L.H[k] = k1_left(g, mbase + k);
        L.E[k] = K1_NEG;            // column b.left takes no horizontal move (fwd2c.h:422)
    }
    L.eopen = 0;
    L.hdiag = k1_left(g, mbase - 1);
}

// One column for one lane.  (h_up, g_up) = Ht(mbase-1, n) and the vertical state G(mbase-1, n) as it
// competed there; first_row: mbase is the first row of the matrix (no vertical move, fwd2c.h:401).
// Returns the 64 direction bits of the 16 cells and the pair to hand to the lane below.
template <int R>
PG_HD unsigned long long k2_lane_step(K2Lane<R>& L, const int* sc, int negv, int h_up, int g_up, bool first_row,
                                      int* h_dn, int* g_dn)
{
    int diag = L.hdiag;
    // vertical state entering row mbase: open from H(mbase-1,n) only if strictly better
    const int x0 = h_up + negv;
    bool gopen = x0 > g_up;
    int g = x0 > g_up ? x0 : g_up;
    if (first_row) { g = K1_NEG; gopen = false; }
    unsigned long long bits = L.eopen;          // bit 3 of every nibble is already in place
    unsigned long long eopen_next = 0;
    int h = h_up, gcur = g_up;
    // The VALUES go through max() only, so that the chain from one row to the next is two operations
    // (h = max(t, g); g' = max(h - v, g)) and everything else -- t = max(diagonal, horizontal), the tie-rule
    // predicates that become direction bits, the horizontal state of the next column -- hangs off that chain
    // (a single warp per scheduler runs these kernels' long pairs: the dependent chain is the step time).
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int d = diag + sc[k];
        const int e = L.E[k];
        const int t = d > e ? d : e;
        const bool e_gt_d = e > d;
        const bool fge = e >= g;                 // F beats G on ties (fwd2c.h:431)
        const bool nd = e_gt_d || g > d;         // a gap state wins only if strictly better (:453)
        h = t > g ? t : g;
        const unsigned src = nd ? (fge ? 2u : 1u) : 0u;
        const unsigned nib = src | (gopen ? 4u : 0u);
        bits |= (unsigned long long)nib << (4 * k);
        const int x = h + negv;
        gcur = g;
        gopen = x > g;                           // next row's vertical state (fwd2c.h:405-408)
        g = x > g ? x : g;
        const bool eo = x > e;                   // next column's horizontal state (:426-429)
        L.E[k] = x > e ? x : e;
        eopen_next |= (unsigned long long)(eo ? 8u : 0u) << (4 * k);
        diag = L.H[k];
        L.H[k] = h;
    }
    L.eopen = eopen_next;
    L.hdiag = h_up;
    *h_dn = h;
    *g_dn = gcur;
    return bits;
}

// ---- direction-bit addressing -------------------------------------------------------------------
// One lane-step owns R / 2 bytes (4 bits per row); with R = 16 that is the 64-bit word above.  Sizes in 64-bit words.
PG_HD long long k2_words_per_pair(int LQ, int LS, int R)
{
    if (LQ <= 0 || LS <= 0) return 0;
    const int rpp = 32 * R;
    const int npass = (LQ + rpp - 1) / rpp;
    const int last_rows = LQ - (npass - 1) * rpp;
    const int last_lanes = (last_rows + R - 1) / R;
    const long long slots = ((long long)(npass - 1) * (LS + 31) + (LS + last_lanes - 1)) * 32;
    return (slots * (R / 2) + 7) / 8;
}

PG_HD unsigned k2_nibble(const unsigned long long* words, int LS, int R, int m, int n)
{
    const int rpp = 32 * R;
    const int pass = m / rpp, rm = m - pass * rpp;
    const int lane = rm / R, k = rm - lane * R;
    const long long slot = ((long long)pass * (LS + 31) + (n + lane)) * 32 + lane;
    const unsigned char b = reinterpret_cast<const unsigned char*>(words)[slot * (R / 2) + (k >> 1)];
    return (unsigned)(b >> (4 * (k & 1))) & 15u;
}

struct K2Rec { int m, n, p; };

// Traceback + record replay for one pair.  `moves` (capacity LQ+LS+2 bytes) and `recs` (capacity
// LQ+LS+4) are scratch.  Writes the corner list in Vmf back-walk order into out (capacity
// LQ+LS+4 entries of 2 ints) with absolute coordinates (window offsets ql, sl added) and returns the
// number of corners.  Move codes: 1 D, 2 V open, 3 V extend, 4 H open, 5 H extend, 6 boundary H,
// 7 boundary V.
// walk the direction bits back from the end cell; returns the number of moves (last move first)
PG_HD int k2_backwalk(const unsigned long long* words, int LQ, int LS, int R, unsigned char* moves)
{
    int nmv = 0;
    int m = LQ - 1, n = LS - 1, state = 0;       // 0 H, 1 G, 2 F
    while (m >= 0 && n >= 0) {
        const unsigned nib = k2_nibble(words, LS, R, m, n);
        if (state == 0) {
            const unsigned src = nib & 3u;
            if (src == 0) { moves[nmv++] = 1; --m; --n; }
            else state = (int)src;
        } else if (state == 1) {
            if (nib & 4u) { moves[nmv++] = 2; state = 0; } else moves[nmv++] = 3;
            --m;
        } else {
            if (nib & 8u) { moves[nmv++] = 4; state = 0; } else moves[nmv++] = 5;
            --n;
        }
    }
    for (; n >= 0; --n) moves[nmv++] = 6;
    for (; m >= 0; --m) moves[nmv++] = 7;
    return nmv;
}

// forward replay of the moves with the reference's record rules, then Vmf::traceback
PG_HD int k2_replay(const unsigned char* moves, int nmv, int LQ, int LS, int ql, int sl, K2Rec* recs, int* out)
{

    // forward replay of the reference's (dir, ptr) bookkeeping along the path
    int nrec = 0;
    recs[nrec].m = ql; recs[nrec].n = sl; recs[nrec].p = -1; ++nrec;    // origin (fwd2c.h:145)
    int hdir = K2_DIAG, hptr = 0;       // H record of the current path cell
    int gdir = 0, gptr = 0;             // running gap-state record
    int cm = -1, cn = -1;               // DP cell of the last move (relative)
    for (int i = nmv - 1; i >= 0; --i) {
        const int mv = moves[i];
        const int nxt = i > 0 ? moves[i - 1] : 0;
        switch (mv) {
        case 6: ++cn; hdir = K2_HORI; break;                                  // boundary row (initB)
        case 7: ++cm; hdir = K2_VERT; break;                                  // boundary column
        case 1:
            ++cm; ++cn;
            hdir = k2_isdiag(hdir) ? K2_DIAG : K2_NEWD;                       // update(), d3 == 0
            if (hdir == K2_NEWD) { recs[nrec].m = cm + ql; recs[nrec].n = cn + sl; recs[nrec].p = hptr; hptr = nrec++; }
            break;
        case 2: case 3:
            ++cm;
            if (mv == 2) { gdir = k2_ishori(hdir) ? K2_NEWV : K2_VERT; gptr = hptr; }   // from H(m-1,n)
            else gdir = K2_VERT;                                                          // from G(m-1,n)
            if (nxt != 3) {             // the run ends here: H(m,n) is a copy of G(m,n)
                hdir = gdir; hptr = gptr;
                if (hdir == K2_NEWV) { recs[nrec].m = cm + ql; recs[nrec].n = cn + sl; recs[nrec].p = hptr; hptr = nrec++; }
            }
            break;
        case 4: case 5:
            ++cn;
            if (mv == 4) { gdir = k2_isvert(hdir) ? K2_NEWH : K2_HORI; gptr = hptr; }
            else gdir = K2_HORI;
            if (nxt != 5) {
                hdir = gdir; hptr = gptr;
                if (hdir == K2_NEWH) { recs[nrec].m = cm + ql; recs[nrec].n = cn + sl; recs[nrec].p = hptr; hptr = nrec++; }
            }
            break;
        }
    }
    // final record (fwd2c.h:476) and Vmf::traceback (vmf.cc:103-119)
    int cnt = 0;
    out[2 * cnt] = LQ + ql; out[2 * cnt + 1] = LS + sl; ++cnt;
    for (int q = hptr; q >= 0; q = recs[q].p) { out[2 * cnt] = recs[q].m; out[2 * cnt + 1] = recs[q].n; ++cnt; }
    return cnt;
}

PG_HD int k2_trace(const unsigned long long* words, int LQ, int LS, int R, int ql, int sl, unsigned char* moves,
                   K2Rec* recs, int* out)
{
    const int nmv = k2_backwalk(words, LQ, LS, R, moves);
    return k2_replay(moves, nmv, LQ, LS, ql, sl, recs, out);
}
