// Device-side gunzip of the hit files (blast_file_processor.py:59-65 opens Blast{i}_{j}.txt.gz through gzip.open;
// DIAMOND writes them compressed, config.json:48-52).  The compressed bytes cross PCIe, the text never does.
//
// One WARP per gzip member.  DEFLATE (RFC 1951) is serial inside a member, so the parallelism comes from members: files
// written as many small members (bgzip / BGZF, `pigz -i`, or this repo's writer: every member carries its compressed
// size in a "BC" extra field, so the host can hop from member to member without inflating) are inflated by thousands of
// warps at once; a classic single-member file is one warp's work (correct, but only as parallel as there are files).
// Every lane runs the same decoder on the same bits (warp-uniform control flow, shared-memory tables are broadcast
// reads); literals are stored by one lane, matches are copied by all lanes (out[pos + i] = out[pos - dist + i % dist]).
// The CRC-32 of the trailer is not verified (structural errors and a wrong ISIZE are).
#pragma once
#ifndef OFG_HOST_SHIM   // tools/micro/inflate_host_check.cpp compiles this file for the CPU with a one-lane "warp"
#include "ofg_common.cuh"
#endif

struct GzMember {
    u64 in_off;     // first byte of the deflate stream in the compressed arena
    u64 in_len;     // bytes of the deflate stream
    u64 out_off;    // first byte of the member's text in the text arena
    u64 out_len;    // ISIZE of the trailer (bytes this member inflates to)
    int file;       // index of the hit file (error reporting)
    int pad;
};

constexpr int GZ_TB = 10;                 // bits of the primary literal/length table
constexpr int GZ_TD = 9;                  // bits of the primary distance table
constexpr int GZ_WARPS = 4;
constexpr int GZ_RING = 2048;             // output ring per warp: matches up to GZ_RING - 258 bytes back never touch global memory
constexpr int GZ_FLUSH = 512;             // bytes written back per flush (coalesced)
#ifdef OFG_HOST_SHIM
constexpr int GZ_LANES = 1;
#else
constexpr int GZ_LANES = 32;
#endif

struct GzWarp {                           // shared memory of one warp (about 6.3 KB)
    unsigned short lit[1 << GZ_TB];       // (symbol << 4) | code length; 0: code longer than GZ_TB bits
    unsigned short dist[1 << GZ_TD];
    unsigned short lsym[288], dsym[32];   // symbols ordered by (code length, symbol): canonical decode of long codes
    unsigned short lcnt[16], dcnt[16];    // codes per length
    unsigned char len[320];               // code lengths while a dynamic header is read
    unsigned short offs[16];              // scratch of gz_build
    u32 next[16];
    int ok;
    u32 win[GZ_RING / 4];                 // the last GZ_RING bytes of the member's output (ring by output position)
};

struct GzBits {
    const uint8_t* in;    // deflate stream
    u64 n;                // its length
    u64 pos;              // next byte to load
    u64 buf;
    int cnt;
    __device__ __forceinline__ void refill()
    {
        // top the buffer up to at least 32 valid bits with one (unaligned) 4-byte read; bytes past the end read as zero
        if (cnt <= 32) {
            const uintptr_t a = (uintptr_t)(in + pos);
            const u32* w = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
            const int sh = (int)(a & 3) * 8;
            u32 v = 0u;
            if (pos < n) {   // never read beyond the stream (a corrupt stream keeps asking for bits: it gets zeros)
                v = __funnelshift_r(w[0], w[1], sh);
                if (pos + 4 > n) v &= 0xffffffffu >> (8 * (int)(pos + 4 - n));
            }
            buf |= (u64)v << cnt;
            cnt += 32;
            pos += 4;
        }
    }
    __device__ __forceinline__ u32 peek(int k) const { return (u32)buf & ((1u << k) - 1u); }
    __device__ __forceinline__ void drop(int k) { buf >>= k; cnt -= k; }
    __device__ __forceinline__ u32 get(int k) { const u32 v = peek(k); drop(k); return v; }
    __device__ __forceinline__ bool overrun() const { return pos > n + 8 && cnt < 0; }
};

__device__ __forceinline__ u32 gz_rev(u32 code, int len) { return __brev(code) >> (32 - len); }

// Canonical Huffman tables from code lengths: lane 0 builds them in shared memory, the other lanes wait.  Returns false for
// an over-subscribed code.
__device__ bool gz_build(GzWarp& sw, const unsigned char* len, int n, unsigned short* tab, int tb, unsigned short* sym, unsigned short* cnt, int lane)
{
    __syncwarp();
    if (lane == 0) {
        for (int l = 0; l < 16; l++) cnt[l] = 0;
        for (int s = 0; s < n; s++) cnt[len[s]]++;
        cnt[0] = 0;
        int left = 1;
        bool ok = true;
        for (int l = 1; l < 16; l++) { left = (left << 1) - (int)cnt[l]; if (left < 0) ok = false; }
        sw.ok = ok ? 1 : 0;
        if (ok) {
            sw.offs[1] = 0; sw.next[1] = 0;
            for (int l = 1; l < 15; l++) { sw.offs[l + 1] = (unsigned short)(sw.offs[l] + cnt[l]); sw.next[l + 1] = (sw.next[l] + cnt[l]) << 1; }
            for (int i = 0; i < (1 << tb); i++) tab[i] = 0;
            for (int s = 0; s < n; s++) {
                const int l = len[s];
                if (!l) continue;
                sym[sw.offs[l]++] = (unsigned short)s;
                const u32 code = sw.next[l]++;
                if (l <= tb) {
                    const u32 r = gz_rev(code, l);
                    for (u32 i = r; i < (1u << tb); i += 1u << l) tab[i] = (unsigned short)((s << 4) | l);
                }
            }
        }
    }
    __syncwarp();
    return sw.ok != 0;
}

// One symbol: primary table, or the canonical bit-by-bit walk for codes longer than the table (rare)
__device__ __forceinline__ int gz_decode(GzBits& b, const unsigned short* tab, int tb, const unsigned short* sym, const unsigned short* cnt)
{
    const u32 e = tab[b.peek(tb)];
    if (e) { b.drop((int)(e & 15u)); return (int)(e >> 4); }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        code |= (int)b.get(1);
        const int c = cnt[l];
        if (code - c < first) return sym[index + (code - first)];
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

__device__ __constant__ unsigned short GZ_LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__device__ __constant__ unsigned char GZ_LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__device__ __constant__ unsigned short GZ_DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__device__ __constant__ unsigned char GZ_DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__device__ __constant__ unsigned char GZ_CLORD[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// err: min over failing members of (member index << 8 | kind); kinds: 1 bad block type / header, 2 bad code, 3 distance too far,
// 4 output overrun, 5 output short of ISIZE, 6 stored block length check
//
// Output path: every byte goes into a shared-memory ring indexed by its output position; whole GZ_FLUSH-byte pieces are
// written back to global memory with coalesced word stores.  A match whose distance is below GZ_RING - 258 reads the
// ring (a few shared-memory cycles instead of a global round trip per symbol); longer distances read bytes that were
// written back long ago.
__global__ void __launch_bounds__(GZ_WARPS * 32) inflate_kernel(const uint8_t* comp, const GzMember* members, int64_t n_members, uint8_t* text,
                                                                 unsigned long long* err)
{
    __shared__ GzWarp s_all[GZ_WARPS];
    GzWarp& sw = s_all[threadIdx.x >> 5];
    uint8_t* const ring = reinterpret_cast<uint8_t*>(sw.win);
    const int lane = (int)lane_id();
    const int64_t n_warps = (int64_t)gridDim.x * GZ_WARPS;
    for (int64_t mi = (int64_t)blockIdx.x * GZ_WARPS + (threadIdx.x >> 5); mi < n_members; mi += n_warps) {
        const GzMember m = members[mi];
        GzBits b;
        b.in = comp + m.in_off; b.n = m.in_len; b.pos = 0; b.buf = 0; b.cnt = 0;
        uint8_t* out = text + m.out_off;   // 4-byte aligned: members start at multiples of 16 of the arena or at BGZF chunk ends
        const u64 head = (u64)((4 - ((uintptr_t)out & 3)) & 3);   // bytes before the first aligned word are stored one by one
        u64 op = 0, flushed = 0;
        int kind = 0;
        bool last = false;
        // ring position = output position + head, so that ring words line up with aligned words of the output
        while (!last && !kind) {
            b.refill();
            last = b.get(1) != 0;
            const u32 type = b.get(2);
            if (type == 0) {
                // stored: skip to the byte boundary, LEN, NLEN, raw bytes
                b.drop(b.cnt & 7);
                b.refill();
                const u32 ln = b.get(16);
                b.refill();
                const u32 nl = b.get(16);
                if ((ln ^ nl) != 0xffffu) { kind = 6; break; }
                const u64 src = b.pos - (u64)(b.cnt >> 3);   // bytes still in the bit buffer belong to the block
                if (src + ln > m.in_len || op + ln > m.out_len) { kind = 4; break; }
                // bytes of the ring that were not written back yet go out one by one, then the block is copied directly and the
                // ring is reloaded with the last bytes of the output
                __syncwarp();
                for (u64 q = flushed + (u64)lane; q < op; q += GZ_LANES) out[q] = ring[(q + head) & (GZ_RING - 1)];
                for (u32 i = lane; i < ln; i += GZ_LANES) out[op + i] = b.in[src + i];
                op += ln;
                __syncwarp();
                const u64 lo = op > (u64)GZ_RING ? op - GZ_RING : 0;
                for (u64 q = lo + (u64)lane; q < op; q += GZ_LANES) ring[(q + head) & (GZ_RING - 1)] = out[q];
                flushed = op;
                b.pos = src + ln; b.buf = 0; b.cnt = 0;
                __syncwarp();
                continue;
            }
            if (type == 3) { kind = 1; break; }
            if (type == 1) {
                for (int s = lane; s < 288; s += GZ_LANES) sw.len[s] = (unsigned char)(s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
                __syncwarp();
                gz_build(sw, sw.len, 288, sw.lit, GZ_TB, sw.lsym, sw.lcnt, lane);
                for (int s = lane; s < 30; s += GZ_LANES) sw.len[s] = 5;
                __syncwarp();
                gz_build(sw, sw.len, 30, sw.dist, GZ_TD, sw.dsym, sw.dcnt, lane);
            } else {
                b.refill();
                const int hlit = (int)b.get(5) + 257, hdist = (int)b.get(5) + 1, hclen = (int)b.get(4) + 4;
                if (hlit > 286 || hdist > 30) { kind = 1; break; }
                __syncwarp();
                for (int s = lane; s < 19; s += GZ_LANES) sw.len[s] = 0;
                __syncwarp();
                for (int i = 0; i < hclen; i++) {
                    b.refill();
                    const u32 v = b.get(3);
                    if (lane == 0) sw.len[GZ_CLORD[i]] = (unsigned char)v;
                }
                // code-length code: reuse the distance table slots for it
                if (!gz_build(sw, sw.len, 19, sw.dist, 7, sw.dsym, sw.dcnt, lane)) { kind = 2; break; }
                int i = 0;
                int prev = 0;
                while (i < hlit + hdist) {
                    b.refill();
                    const int s = gz_decode(b, sw.dist, 7, sw.dsym, sw.dcnt);
                    if (s < 0) { kind = 2; break; }
                    if (s < 16) {
                        if (lane == 0) sw.len[i] = (unsigned char)s;
                        prev = s;
                        i++;
                    } else {
                        int rep, val = 0;
                        if (s == 16) { if (i == 0) { kind = 2; break; } val = prev; rep = 3 + (int)b.get(2); }
                        else if (s == 17) rep = 3 + (int)b.get(3);
                        else rep = 11 + (int)b.get(7);
                        if (i + rep > hlit + hdist) { kind = 2; break; }
                        if (lane == 0) for (int k = 0; k < rep; k++) sw.len[i + k] = (unsigned char)val;
                        if (s != 16) prev = 0;
                        i += rep;
                    }
                }
                if (kind) break;
                __syncwarp();
                // the distance lengths follow the literal/length ones: build the distance table first (gz_build reads them in place)
                if (sw.len[256] == 0) { kind = 2; break; }
                if (!gz_build(sw, sw.len + hlit, hdist, sw.dist, GZ_TD, sw.dsym, sw.dcnt, lane)) { kind = 2; break; }
                if (!gz_build(sw, sw.len, hlit, sw.lit, GZ_TB, sw.lsym, sw.lcnt, lane)) { kind = 2; break; }
            }
            // ---- symbols of the block
            for (;;) {
                // write whole pieces of the ring back while they are at least a maximal match away from being overwritten
                {
                    // aligned window [fa, fb): fa = first unflushed aligned position, in ring coordinates (output position + head)
                    const u64 done = flushed;   // output positions below `done` are in global memory
                    if (op - done >= (u64)(GZ_FLUSH + 8)) {
                        // bytes [done, done + head') one by one until aligned, then words
                        u64 a0 = done;
                        if (((a0 + head) & 3) != 0) {
                            const u64 a1 = a0 + (4 - ((a0 + head) & 3));
                            __syncwarp();
                            for (u64 q = a0 + (u64)lane; q < a1; q += GZ_LANES) out[q] = ring[(q + head) & (GZ_RING - 1)];
                            a0 = a1;
                        }
                        const u64 a2 = a0 + GZ_FLUSH;
                        __syncwarp();
                        for (u64 q = a0 + 4 * (u64)lane; q < a2; q += 4 * GZ_LANES)
                            *reinterpret_cast<u32*>(out + q) = sw.win[((q + head) & (GZ_RING - 1)) >> 2];
                        flushed = a2;
                        __syncwarp();
                    }
                }
                b.refill();
                int s = gz_decode(b, sw.lit, GZ_TB, sw.lsym, sw.lcnt);
                if (s < 256) {
                    if (s < 0) { kind = 2; break; }
                    if (op >= m.out_len) { kind = 4; break; }
                    if (lane == 0) ring[(op + head) & (GZ_RING - 1)] = (uint8_t)s;
                    op++;
                    continue;
                }
                if (s == 256) break;
                s -= 257;
                if (s >= 29) { kind = 2; break; }
                const int ln = (int)GZ_LBASE[s] + (int)b.get((int)GZ_LEXT[s]);
                b.refill();
                const int ds = gz_decode(b, sw.dist, GZ_TD, sw.dsym, sw.dcnt);
                if (ds < 0 || ds >= 30) { kind = 2; break; }
                const u64 dist = (u64)GZ_DBASE[ds] + (u64)b.get((int)GZ_DEXT[ds]);
                if (dist > op) { kind = 3; break; }
                if (op + (u64)ln > m.out_len) { kind = 4; break; }
                __syncwarp();   // earlier stores of this warp are visible to the loads below
                if (dist <= (u64)(GZ_RING - 264)) {
                    // source bytes are still in the ring (period-dist pattern when the match overlaps itself)
                    const u32 sb = (u32)(op + head - dist);
                    if (dist >= (u64)ln) {
                        for (int i = lane; i < ln; i += GZ_LANES) ring[(u32)(op + head + i) & (GZ_RING - 1)] = ring[(sb + (u32)i) & (GZ_RING - 1)];
                    } else {
                        const u32 d32 = (u32)dist;
                        uint8_t v[9];   // read everything first: with GZ_LANES lanes a source byte may be a destination of this very match
#pragma unroll
                        for (int k = 0; k < 9; k++) { const int i = lane + k * GZ_LANES; v[k] = (GZ_LANES > 1 && i < ln) ? ring[(sb + (u32)i % d32) & (GZ_RING - 1)] : 0; }
                        if (GZ_LANES > 1) {
#pragma unroll
                            for (int k = 0; k < 9; k++) { const int i = lane + k * GZ_LANES; if (i < ln) ring[(u32)(op + head + i) & (GZ_RING - 1)] = v[k]; }
                        } else {
                            for (int i = 0; i < ln; i++) ring[(u32)(op + head + i) & (GZ_RING - 1)] = ring[(sb + (u32)i % d32) & (GZ_RING - 1)];
                        }
                    }
                } else {
                    // far back: those bytes were written to global memory long ago (flushed > op - GZ_FLUSH - 270 > op - dist + ln)
                    const uint8_t* src = out + op - dist;
                    for (int i = lane; i < ln; i += GZ_LANES) ring[(u32)(op + head + i) & (GZ_RING - 1)] = src[i];
                }
                op += (u64)ln;
                __syncwarp();
            }
        }
        // the rest of the ring
        __syncwarp();
        if (!kind) for (u64 q = flushed + (u64)lane; q < op; q += GZ_LANES) out[q] = ring[(q + head) & (GZ_RING - 1)];
        if (!kind && op != m.out_len) kind = 5;
        if (kind && lane == 0) atomicMin(err, ((unsigned long long)mi << 8) | (unsigned long long)kind);
        __syncwarp();
    }
}
