// ref_wrap.cpp -- C-ABI wrapper around the UNMODIFIED reference vec_query().
//
// TEST INFRASTRUCTURE ONLY (see oracle/hvs_oracle.c header).  This file contains no
// reference code: it #includes the reference headers where they lie under
// /root/reference/include (passed with -I by oracle/Makefile) and is compiled once per
// -DIMPL value into oracle/_ref/libref_{baseline,optimized,parallel}.so, mirroring the
// compile-time switch of src/test.cpp:6-13.  The flat-buffer -> vector<vector<float>>
// conversion reproduces what ReadBin (include/io.h:111-136) hands to vec_query.
#include "io.h"
#if IMPL == 2
#include "optimized.hpp"
#elif IMPL == 3
#include "optimized_parallel.hpp"
#else
#include "baseline.hpp"
#endif
#include <chrono>
#include <sstream>

extern "C" {

// Returns wall seconds spent inside vec_query only (the region src/test.cpp:82-88 times),
// or a negative value on bad arguments.  out_ids: nq x 100 uint32.
double ref_vec_query(const float *nodes_flat, uint32_t n, const float *queries_flat, uint32_t nq,
                     float sample_proportion, uint32_t *out_ids)
{
    if (!nodes_flat || !queries_flat || !out_ids || n < 100) return -1.0;
    std::vector<std::vector<float>> nodes(n), queries(nq);
    for (uint32_t i = 0; i < n; ++i)
        nodes[i].assign(nodes_flat + (size_t)i * 102, nodes_flat + (size_t)(i + 1) * 102);
    for (uint32_t i = 0; i < nq; ++i)
        queries[i].assign(queries_flat + (size_t)i * 104, queries_flat + (size_t)(i + 1) * 104);
    std::vector<std::vector<uint32_t>> knn_results;
    // vec_query prints progress to std::cout; silence it for library use
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    auto t0 = std::chrono::steady_clock::now();
    vec_query(nodes, queries, sample_proportion, knn_results);
    auto t1 = std::chrono::steady_clock::now();
    std::cout.rdbuf(old);
    if (knn_results.size() != nq) return -2.0;
    for (uint32_t i = 0; i < nq; ++i)
        for (int k = 0; k < 100; ++k) out_ids[(size_t)i * 100 + k] = knn_results[i][k];
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_impl(void) { return IMPL; }

// The same with the data set converted ONCE (10^7 rows = 10^7 heap blocks, as ReadBin builds them) and kept for many
// vec_query calls.  ref_query_loaded may be called from several host threads at once for IMPL=1/2 (vec_query only reads
// `nodes`; its PERF_DBG accumulators are plain globals whose values nobody uses here); std::cout is silenced for the
// whole session, not per call, so that concurrent calls do not fight over the stream buffer.
namespace {
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
std::vector<std::vector<float>> g_nodes;
NullBuf g_null;
std::streambuf *g_old_cout = nullptr, *g_old_cerr = nullptr;
}

int ref_set_nodes(const float *nodes_flat, uint32_t n)
{
    if (!nodes_flat || n < 100) return -1;
    g_nodes.assign(n, std::vector<float>());
    for (uint32_t i = 0; i < n; ++i)
        g_nodes[i].assign(nodes_flat + (size_t)i * 102, nodes_flat + (size_t)(i + 1) * 102);
    if (!g_old_cout) g_old_cout = std::cout.rdbuf(&g_null);
    if (!g_old_cerr) g_old_cerr = std::cerr.rdbuf(&g_null);      // the PERF_DBG phase timers (baseline.hpp:179-189)
    return 0;
}

void ref_clear_nodes(void)
{
    std::vector<std::vector<float>>().swap(g_nodes);
    if (g_old_cout) { std::cout.rdbuf(g_old_cout); g_old_cout = nullptr; }
    if (g_old_cerr) { std::cerr.rdbuf(g_old_cerr); g_old_cerr = nullptr; }
}

double ref_query_loaded(const float *queries_flat, uint32_t nq, float sample_proportion, uint32_t *out_ids)
{
    if (!queries_flat || !out_ids || g_nodes.size() < 100) return -1.0;
    std::vector<std::vector<float>> queries(nq);
    for (uint32_t i = 0; i < nq; ++i)
        queries[i].assign(queries_flat + (size_t)i * 104, queries_flat + (size_t)(i + 1) * 104);
    std::vector<std::vector<uint32_t>> knn_results;
    auto t0 = std::chrono::steady_clock::now();
    vec_query(g_nodes, queries, sample_proportion, knn_results);
    auto t1 = std::chrono::steady_clock::now();
    if (knn_results.size() != nq) return -2.0;
    for (uint32_t i = 0; i < nq; ++i)
        for (int k = 0; k < 100; ++k) out_ids[(size_t)i * 100 + k] = knn_results[i][k];
    return std::chrono::duration<double>(t1 - t0).count();
}

}
