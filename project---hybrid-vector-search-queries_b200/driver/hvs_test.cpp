// hvs_test.cpp -- stand-alone driver with the reference driver's command line and file formats
// (src/test.cpp:20-112): hvs_test.out [source_path] [query_path] [output_path].
//   source : uint32 N, then N x 102 float32 [C, T, x0..x99]       (what ReadBin(path,102,..) reads, include/io.h:111-136)
//   query  : uint32 M, then M x 104 float32 [type, v, l, r, q..]  (ReadBin(path,104,..))
//   output : M x 100 uint32, no header                            (SaveKNN, include/io.h:23-36)
//   output.dist : uint32 M, then M x 100 float32                  (SaveKNNFull, include/io.h:50-78)
// Inside the reference tree the same thing is `-DIMPL=4` in src/test.cpp with the reference's own
// io.h; this file only exists so that the engine can be driven without it.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "hvs_vec_query.hpp"

static bool read_bin(const std::string &path, size_t dims, std::vector<std::vector<float>> &rows)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    uint32_t n = 0;
    f.read(reinterpret_cast<char *>(&n), 4);
    rows.resize(n);
    std::vector<float> buf(dims);
    for (uint32_t i = 0; i < n; ++i) {
        f.read(reinterpret_cast<char *>(buf.data()), (std::streamsize)(dims * 4));
        if (!f) return false;
        rows[i] = buf;
    }
    return true;
}

int main(int argc, char **argv)
{
    std::string source = "dummy-data.bin", query = "dummy-queries.bin", out = "output.bin";
    if (argc > 1) source = argv[1];
    if (argc > 2) query = argv[2];
    if (argc > 3) out = argv[3];
    if (argc > 4) { std::cerr << "usage: " << argv[0] << " [source_path] [query_path] [output_path]\n"; return 1; }
    std::vector<std::vector<float>> nodes, queries;
    if (!read_bin(source, HVS_DATA_ROW, nodes)) { std::cerr << "cannot read " << source << "\n"; return 1; }
    if (!read_bin(query, HVS_QUERY_ROW, queries)) { std::cerr << "cannot read " << query << "\n"; return 1; }
    std::cout << "# data points:  " << nodes.size() << "\n# queries:      " << queries.size() << std::endl;
    std::vector<std::vector<uint32_t>> knn;
    std::vector<float> dist;                       // the `.dist` table comes out of the same solve (hvs_solve_full): one index build
    hvs_shim::want_dist = &dist;
    auto t0 = std::chrono::steady_clock::now();
    vec_query(nodes, queries, 1.0f, knn);
    auto t1 = std::chrono::steady_clock::now();
    std::cerr << "Vector Search took " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " ms\n";
    {
        std::ofstream f(out, std::ios::binary);
        for (auto &r : knn) f.write(reinterpret_cast<const char *>(r.data()), (std::streamsize)(r.size() * 4));
    }
    {   // .dist side file: uint32 M, then M x 100 float32 (include/io.h:50-78)
        std::ofstream f(out + ".dist", std::ios::binary);
        uint32_t m = (uint32_t)queries.size();
        f.write(reinterpret_cast<const char *>(&m), 4);
        f.write(reinterpret_cast<const char *>(dist.data()), (std::streamsize)(dist.size() * 4));
    }
    return 0;
}
