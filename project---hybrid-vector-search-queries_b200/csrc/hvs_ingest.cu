// hvs_ingest.cu -- host ingest for the indexing phase: D reaches the GPU through two pinned staging
// buffers that are filled by host threads while the previous chunk is on the wire.
//
// Replaces the tail of the reference's load path for this engine: ReadBin (include/io.h:111-136)
// materialises D as N separate heap blocks of 408 bytes (std::vector<std::vector<float>>), which is
// what vec_query receives.  Three sources, one streaming core:
//   hvs_index_build            one row-major host buffer (pageable or pinned)
//   hvs_index_build_rows       N row pointers -- the nested-vector layout itself, no intermediate copy
//   hvs_index_build_from_file  the D file (uint32 N, then N x 102 float32; README.md:32-44), read with
//                              pread straight into the pinned buffers
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "hvs_engine.h"

using namespace hvs;

namespace {

constexpr size_t CHUNK = 64u << 20;     // bytes per staging buffer

// fill(dst, byte_offset, bytes) must write `bytes` bytes of the row-major image of D starting at byte_offset
int stream_to_device(hvs_engine *e, void *dev, size_t total, const std::function<bool(char *, size_t, size_t)> &fill)
{
    cudaError_t c = e->h_ingest[0].ensure(CHUNK);
    if (c == cudaSuccess) c = e->h_ingest[1].ensure(CHUNK);
    if (c != cudaSuccess) { cudaGetLastError(); e->err = "ingest: pinned staging allocation failed"; return HVS_ERR_NOMEM; }
    cudaEvent_t done[2];
    cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming);
    int rc = HVS_OK;
    size_t off = 0;
    for (int i = 0; off < total; ++i, off += CHUNK) {
        const int b = i & 1;
        const size_t bytes = std::min(CHUNK, total - off);
        if (i >= 2) cudaEventSynchronize(done[b]);                  // the copy that last used this buffer has finished
        if (!fill(e->h_ingest[b].as<char>(), off, bytes)) { if (e->err.empty()) e->err = "ingest: source read failed"; rc = HVS_ERR_INVALID; break; }
        c = cudaMemcpyAsync((char *)dev + off, e->h_ingest[b].p, bytes, cudaMemcpyHostToDevice, e->stream);
        if (c != cudaSuccess) { e->err = std::string("ingest: H2D: ") + cudaGetErrorString(c); rc = HVS_ERR_CUDA; break; }
        cudaEventRecord(done[b], e->stream);
    }
    c = cudaStreamSynchronize(e->stream);
    if (rc == HVS_OK && c != cudaSuccess) { e->err = std::string("ingest: ") + cudaGetErrorString(c); rc = HVS_ERR_CUDA; }
    cudaEventDestroy(done[0]);
    cudaEventDestroy(done[1]);
    return rc;
}

template <class F>
void host_parallel(size_t n, F &&f)                                  // f(lo, hi) over [0, n) on a few threads
{
    unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 8u);
    if (n < (1u << 16)) nt = 1;
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(f, n * t / nt, n * (t + 1) / nt);
    f((size_t)0, n / nt);
    for (auto &t : th) t.join();
}

int check_args(hvs_engine *e, const void *src, uint32_t n, float sample_proportion)
{
    if (!e) return HVS_ERR_INVALID;
    e->err.clear();
    if (!src) { e->err = "hvs_index_build: source is NULL"; return HVS_ERR_INVALID; }
    if (n < HVS_K) {
        e->err = "hvs_index_build: n < 100 (the reference's pad rule reads nodes[n-s], include/baseline.hpp:138-147)";
        return HVS_ERR_INVALID;
    }
    if (!(sample_proportion >= 0.f)) { e->err = "hvs_index_build: sample_proportion must be >= 0"; return HVS_ERR_INVALID; }
    if (cudaSetDevice(e->device) != cudaSuccess) { e->err = "cudaSetDevice failed"; return HVS_ERR_CUDA; }
    return HVS_OK;
}

int finish(hvs_engine *e, DevBuf &rows, int rc, uint32_t n, float sample_proportion)
{
    if (rc == HVS_OK) rc = hvs_index_build_device(e, rows.as<float>(), n, sample_proportion);
    rows.release();
    return rc;
}

}  // namespace

extern "C" int hvs_index_build(hvs_engine *e, const float *rows_host, uint32_t n, float sample_proportion)
{
    int rc = check_args(e, rows_host, n, sample_proportion);
    if (rc) return rc;
    DevBuf rows;
    const size_t total = (size_t)n * DROW * 4;
    if (rows.ensure(total) != cudaSuccess) { cudaGetLastError(); e->err = "hvs_index_build: device allocation for D failed"; return HVS_ERR_NOMEM; }
    const char *src = reinterpret_cast<const char *>(rows_host);
    rc = stream_to_device(e, rows.p, total, [&](char *dst, size_t off, size_t bytes) {
        host_parallel(bytes, [&](size_t lo, size_t hi) { std::memcpy(dst + lo, src + off + lo, hi - lo); });
        return true;
    });
    return finish(e, rows, rc, n, sample_proportion);
}

extern "C" int hvs_index_build_rows(hvs_engine *e, const float *const *row_ptrs, uint32_t n, float sample_proportion)
{
    int rc = check_args(e, row_ptrs, n, sample_proportion);
    if (rc) return rc;
    DevBuf rows;
    constexpr size_t RB = (size_t)DROW * 4;
    const size_t total = (size_t)n * RB;
    if (rows.ensure(total) != cudaSuccess) { cudaGetLastError(); e->err = "hvs_index_build_rows: device allocation for D failed"; return HVS_ERR_NOMEM; }
    static_assert(CHUNK % 8 == 0, "chunk arithmetic");
    rc = stream_to_device(e, rows.p, total, [&](char *dst, size_t off, size_t bytes) {
        // a chunk may start and end inside a row: copy the covered part of every row it touches
        const size_t r0 = off / RB, r1 = (off + bytes + RB - 1) / RB;
        bool ok = true;
        host_parallel(r1 - r0, [&](size_t lo, size_t hi) {
            for (size_t r = r0 + lo; r < r0 + hi; ++r) {
                const char *src = reinterpret_cast<const char *>(row_ptrs[r]);
                if (!src) { ok = false; return; }
                const size_t rb = r * RB, a = std::max(rb, off), b = std::min(rb + RB, off + bytes);
                std::memcpy(dst + (a - off), src + (a - rb), b - a);
            }
        });
        if (!ok) e->err = "hvs_index_build_rows: NULL row pointer";
        return ok;
    });
    return finish(e, rows, rc, n, sample_proportion);
}

extern "C" int hvs_index_build_from_file(hvs_engine *e, const char *path, float sample_proportion, uint32_t *out_n)
{
    if (!e) return HVS_ERR_INVALID;
    e->err.clear();
    if (!path) { e->err = "hvs_index_build_from_file: path is NULL"; return HVS_ERR_INVALID; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { e->err = std::string("hvs_index_build_from_file: cannot open ") + path; return HVS_ERR_INVALID; }
    uint32_t n = 0;
    struct stat sb;
    if (pread(fd, &n, 4, 0) != 4 || fstat(fd, &sb) != 0 || (uint64_t)sb.st_size < 4 + (uint64_t)n * DROW * 4) {
        close(fd);
        e->err = std::string("hvs_index_build_from_file: ") + path + " is shorter than its row count says";
        return HVS_ERR_INVALID;
    }
    if (out_n) *out_n = n;
    int rc = check_args(e, path, n, sample_proportion);
    if (rc) { close(fd); return rc; }
    DevBuf rows;
    const size_t total = (size_t)n * DROW * 4;
    if (rows.ensure(total) != cudaSuccess) { cudaGetLastError(); close(fd); e->err = "hvs_index_build_from_file: device allocation for D failed"; return HVS_ERR_NOMEM; }
    rc = stream_to_device(e, rows.p, total, [&](char *dst, size_t off, size_t bytes) {
        bool ok = true;
        host_parallel(bytes, [&](size_t lo, size_t hi) {
            size_t p = lo;
            while (p < hi) {
                const ssize_t got = pread(fd, dst + p, hi - p, (off_t)(4 + off + p));
                if (got <= 0) { ok = false; return; }
                p += (size_t)got;
            }
        });
        return ok;
    });
    close(fd);
    return finish(e, rows, rc, n, sample_proportion);
}
