// trueknn_cli.cpp — the sample's command line over libtrueknn (C++ host side of the drop-in).
//
//   trueknn <file> <n> <dim> <start radius> <k> <output file> [--neighbours <path>] [--device <d>] [--json]
//
// Mirrors samples/s01-trueknn/hostCode.cpp main() (argv contract :66-73): reads the first n points of
// a text file (grammar :83-104, 2-D/3-D handling :114-124), builds the accel ("Build time",
// :201-212), runs the rounds ("True KNN time", :279-347), prints the same three lines and appends
// "Total time" to the output file (:346-356).  What the reference leaves commented out (:312-319) —
// writing `query,neighbourIndex,distance` lines — is available behind --neighbours.
// Unlike the reference it returns a non-zero exit code with a message instead of perror/exit in the
// middle of the run, and rejects k > n-1 (the reference loops forever, :285,321-323).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../include/trueknn.h"

namespace {

// hostCode.cpp:83-104: while (getline && count > 0) { stringstream ss(line); while (ss >> f) { push; count--; if (peek == ',') ignore; } }
bool read_points_text(const std::string& path, long long n, int dim, std::vector<float>& flat) {
  std::ifstream f(path);
  if (!f.is_open()) return false;
  long long owed = n * dim;
  std::string line;
  while (owed > 0 && std::getline(f, line)) {
    const char* p = line.c_str();
    const char* end = p + line.size();
    while (p < end) {
      while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) ++p;
      if (p >= end) break;
      char* q = nullptr;
      const float v = std::strtof(p, &q);
      if (q == p) break;  // operator>> fails: the rest of the line is dropped
      flat.push_back(v);
      --owed;
      p = q;
      if (p < end && *p == ',') ++p;
    }
  }
  return true;
}

bool read_points_f32(const std::string& path, long long n, int dim, std::vector<float>& flat) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) return false;
  flat.resize((size_t)n * dim);
  f.read(reinterpret_cast<char*>(flat.data()), (std::streamsize)(flat.size() * sizeof(float)));
  flat.resize((size_t)(f.gcount() / (std::streamsize)sizeof(float)) / dim * dim);
  return true;
}

bool ends_with(const std::string& s, const char* suf) {
  const size_t m = std::strlen(suf);
  return s.size() >= m && s.compare(s.size() - m, m, suf) == 0;
}

}  // namespace

int main(int ac, char** av) {
  std::vector<std::string> pos;
  std::string neigh_path;
  int device = 0;
  bool json = false;
  for (int i = 1; i < ac; ++i) {
    const std::string a = av[i];
    if (a == "--neighbours" || a == "--neighbors") { if (++i < ac) neigh_path = av[i]; }
    else if (a == "--device") { if (++i < ac) device = std::atoi(av[i]); }
    else if (a == "--json") json = true;
    else pos.push_back(a);
  }
  if (pos.size() != 6) {
    std::fprintf(stderr, "usage: %s <file> <n> <dim> <start radius> <k> <output file> [--neighbours <path>] [--device <d>] [--json]\n", av[0]);
    return 64;
  }
  const std::string path = pos[0], outfile = pos[5];
  const long long n = std::atoll(pos[1].c_str());
  const int dim = std::atoi(pos[2].c_str());
  const float radius = (float)std::atof(pos[3].c_str());
  const int k = std::atoi(pos[4].c_str());
  if (dim != 2 && dim != 3) { std::fprintf(stderr, "dimension must be 2 or 3\n"); return 65; }

  std::vector<float> flat;
  const bool ok = ends_with(path, ".f32") ? read_points_f32(path, n, dim, flat) : read_points_text(path, n, dim, flat);
  if (!ok) { std::perror("Error open"); return 66; }
  if (flat.size() % (size_t)dim) {
    std::fprintf(stderr, "%zu floats is not a multiple of dim=%d (the reference throws std::out_of_range here)\n", flat.size(), dim);
    return 65;
  }
  const uint64_t np = flat.size() / (size_t)dim;
  std::cout << " num spheres: " << np << "\n";

  tknn_ctx* ctx = nullptr;
  int rc = tknn_create(device, &ctx);
  if (rc != TKNN_OK) { std::fprintf(stderr, "tknn_create failed (%d): a CUDA sm_100 device is required\n", rc); return 70; }

  auto t0 = std::chrono::steady_clock::now();
  rc = tknn_build(ctx, flat.data(), np, dim, dim);
  auto t1 = std::chrono::steady_clock::now();
  if (rc != TKNN_OK) { std::fprintf(stderr, "tknn_build: %s\n", tknn_last_error(ctx)); tknn_destroy(ctx); return 71; }
  const double build_s = std::chrono::duration<double>(t1 - t0).count();
  std::cout << "Build time: " << build_s << '\n';

  std::vector<int32_t> idx((size_t)np * k);
  std::vector<float> dist((size_t)np * k);
  auto t2 = std::chrono::steady_clock::now();
  rc = tknn_search(ctx, k, radius, idx.data(), dist.data());
  auto t3 = std::chrono::steady_clock::now();
  if (rc != TKNN_OK) { std::fprintf(stderr, "tknn_search: %s\n", tknn_last_error(ctx)); tknn_destroy(ctx); return 72; }
  const double knn_s = std::chrono::duration<double>(t3 - t2).count();
  tknn_stats st;
  tknn_get_stats(ctx, &st);
  float r = st.start_radius;
  for (int i = 0; i < st.rounds && i < TKNN_MAX_ROUNDS; ++i, r *= 2) {
    std::cout << "Round: " << (i + 1) << " Radius = " << r << " Queries = " << st.round_queries[i]
              << " Time: " << st.round_ms[i] / 1000.0 << " seconds\n";
  }
  std::cout << "True KNN time: " << knn_s << " seconds." << std::endl;
  const double tot = build_s + knn_s;
  std::cout << "Total time: " << tot << '\n';

  std::ofstream out(outfile, std::ios::app);
  if (!out.is_open()) { std::perror("Error open"); tknn_destroy(ctx); return 73; }
  out << tot << std::endl;

  if (!neigh_path.empty()) {
    std::FILE* f = std::fopen(neigh_path.c_str(), "w");
    if (!f) { std::perror("Error open"); tknn_destroy(ctx); return 73; }
    for (uint64_t j = 0; j < np; ++j)
      for (int i = 0; i < k; ++i) std::fprintf(f, "%llu,%d,%.9g\n", (unsigned long long)j, idx[j * k + i], dist[j * k + i]);
    std::fclose(f);
  }
  if (json) {
    std::printf("{\"n\": %llu, \"k\": %d, \"rounds\": %d, \"build_ms\": %.4f, \"search_ms\": %.4f, \"start_radius\": %.9g, "
                "\"queries_per_s\": %.1f}\n",
                (unsigned long long)np, k, st.rounds, st.build_ms, st.search_ms, st.start_radius,
                st.search_ms > 0 ? np / (st.search_ms * 1e-3) : 0.0);
  }
  tknn_destroy(ctx);
  return 0;
}
