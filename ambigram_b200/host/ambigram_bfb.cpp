// ambigram_bfb.cpp -- `--op bfb` host driver on top of libambigram_b200 (C ABI).
//
// A from-scratch C++ host that mirrors what the reference's CLI does around the hot path
// (localhap.cpp:49-389), so that the library can be exercised end to end from `.lh` text to the printed
// BFB path without the reference tree:
//   .lh reader                      Graph::readGraph            src/Graph.cpp:109-237
//   depth -> copy number            calculateHapDepth/CopyNum   src/Graph.cpp:312-405
//   .juncs reader                   readComponents              src/LocalGenomicMap.cpp:5096-5156
//   fold-back copy numbers          getJuncCN                   src/LocalGenomicMap.cpp:3989-4050
//   indel pre-adjustment            getIndelBias                src/LocalGenomicMap.cpp:3699-3744
//   Stage A                         ambi_bfb_solve[_batch]      (replaces localhap.cpp:173-212)
//   Stage B                         ambi_bfb_construct_dag + ambi_bfb_first_valid_order (localhap.cpp:235-261)
//   imperfect fold-backs, expansion ambi_bfb_imperfect_fbi, ambi_bfb_expand_path
//   indel post-splice               indelBFB                    src/LocalGenomicMap.cpp:3746-3837
//   printing                        printBFB                    src/LocalGenomicMap.cpp:3411-3429
// Translocation modes (PROP I1/I2/C1/C2) are out of scope.
//
//   ambigram_bfb --op bfb --in_lh X.lh [--juncdb F --junc_info true] [--reversed] [--all] [--lp_prefix P]
//   ambigram_bfb --op bfb --in_list LIST     one .lh path per line: all partitions go through one
//                                            ambi_bfb_solve_batch / first_valid_order_batch call
#include "../../include/ambigram_b200.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

struct Seg { int id = 0, chr_id = 0, start = 0, end = 0; std::string chrom; double cov = 0, cn = 0; };
struct Junc { int s = 0, t = 0; char sd = '+', td = '+'; double cov = 0, cn = 0; };

struct LhGraph {
    std::vector<Seg> segs;                 // in file order; ids are 1..N in every supported input
    std::vector<Junc> juncs;
    std::vector<int> sources, sinks;
    double avg_cov_raw = -1, purity = -1, avg_tumor_ploidy = -1, avg_ploidy = 0;
    int virus_start = 1 << 30;
    Seg *seg(int id) { for (Seg &s : segs) if (s.id == id) return &s; return nullptr; }
};

static std::vector<std::string> split(const std::string &s, char c)
{
    std::vector<std::string> out; std::string cur; std::istringstream is(s);
    while (std::getline(is, cur, c)) out.push_back(cur);
    return out;
}

static bool parse_lh(const std::string &fn, LhGraph &g)
{
    std::ifstream in(fn);
    if (!in) { std::cerr << "Cannot open file " << fn << std::endl; return false; }
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string key; ss >> key;
        if (key.empty() || key[0] == '#') continue;
        if (key == "AVG_WHOLE_HOST_DP") ss >> g.avg_cov_raw;
        else if (key == "PURITY") ss >> g.purity;
        else if (key == "AVG_TUMOR_PLOIDY") ss >> g.avg_tumor_ploidy;
        else if (key == "AVG_PLOIDY") ss >> g.avg_ploidy;
        else if (key == "VIRUS_START") ss >> g.virus_start;
        else if (key == "SOURCE" || key == "SINK") {
            std::string v; ss >> v;
            for (const std::string &t : split(v, ',')) (key == "SOURCE" ? g.sources : g.sinks).push_back(atoi(t.c_str()));
        } else if (key == "SEG") {
            std::string node; Seg s; ss >> node >> s.cov >> s.cn;
            std::vector<std::string> f = split(node, ':');
            if (f.size() < 5) { std::cerr << "malformed SEG line: " << line << std::endl; return false; }
            s.id = atoi(f[1].c_str()); s.chrom = f[2]; s.start = atoi(f[3].c_str()); s.end = atoi(f[4].c_str());
            s.cov = std::max(s.cov, 0.0);
            for (size_t i = 0; i < g.sources.size() && i < g.sinks.size(); ++i)
                if (g.sources[i] <= s.id && s.id <= g.sinks[i]) s.chr_id = (int)i;
            g.segs.push_back(s);
        } else if (key == "JUNC") {
            std::string a, b; Junc j; ss >> a >> b >> j.cov >> j.cn;
            if (j.cov <= 0 && j.cn <= 0) continue;
            std::vector<std::string> fa = split(a, ':'), fb = split(b, ':');
            if (fa.size() < 3 || fb.size() < 3) { std::cerr << "malformed JUNC line: " << line << std::endl; return false; }
            j.s = atoi(fa[1].c_str()); j.sd = fa[2][0]; j.t = atoi(fb[1].c_str()); j.td = fb[2][0];
            g.juncs.push_back(j);
        }
    }
    if (g.sources.size() != g.sinks.size()) { std::cerr << "SOURCE/SINK count mismatch" << std::endl; return false; }
    // depth -> CN only where no positive CN is given (Graph.cpp:373, 390)
    bool need = false;
    for (const Seg &s : g.segs) need |= s.cn <= 0;
    for (const Junc &j : g.juncs) need |= j.cn <= 0;
    if (need) {
        // calculateHapDepth (Graph.cpp:312-367; Graph(const char*) starts with mAvgPloidy = 0, so its "AVG_PLOIDY
        // given" branch always runs): the normal-cell share of the depth, ratio = 1 - pT / (pT + 2(1-p)), is
        // subtracted before dividing by the haploid depth; a given AVG_PLOIDY is kept only within 0.1 of pT + 2(1-p)
        double ap = g.avg_ploidy, ratio = 0;
        if (g.avg_tumor_ploidy >= 0 && g.purity >= 0) {
            const double tumour = g.purity * g.avg_tumor_ploidy, computed = tumour + (1 - g.purity) * 2;
            ratio = 1 - tumour / computed;
            if (std::fabs(ap - computed) > 0.1) ap = computed;
        }
        const double hdp = g.avg_cov_raw * g.purity / ap;
        for (Seg &s : g.segs) if (s.cn <= 0)
            s.cn = std::max(s.id >= g.virus_start ? s.cov / g.avg_cov_raw * 2 : (s.cov - g.avg_cov_raw * ratio) / hdp, 0.0);
        for (Junc &j : g.juncs) if (j.cn <= 0) j.cn = std::max((j.cov - g.avg_cov_raw * ratio) / hdp, 0.0);
    }
    return true;
}

// readComponents (LGM:5096-5156)
static std::vector<std::vector<int>> read_components(LhGraph &g, const std::string &fn)
{
    std::vector<std::vector<int>> res;
    if (fn.empty()) return res;
    std::ifstream in(fn);
    std::string line;
    auto part = [&](int id) { Seg *s = g.seg(id); return s ? s->chr_id : -1; };
    auto comp = [](char d) { return d == '+' ? '-' : '+'; };
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::vector<int> segs; std::vector<char> sign; std::string t;
        while (ss >> t) { segs.push_back(atoi(t.substr(0, t.size() - 1).c_str())); sign.push_back(t.back()); }
        size_t last = 0;
        for (size_t i = 1; i < segs.size(); ++i) {
            if (part(segs[last]) != part(segs[i]) || sign[i - 1] != sign[i]) {
                if (i - last >= 2) { std::vector<int> sub(segs.begin() + last, segs.begin() + i); std::sort(sub.begin(), sub.end()); res.push_back(sub); }
                int s = segs[i - 1], tt = segs[i]; char sd = sign[i - 1], td = sign[i];
                std::cout << s << sd << " -> " << tt << td << std::endl;
                Junc *found = nullptr;
                for (Junc &j : g.juncs)
                    if ((j.s == s && j.sd == sd && j.t == tt && j.td == td) || (j.s == tt && j.sd == comp(td) && j.t == s && j.td == comp(sd))) { found = &j; break; }
                if (!found) { Junc j; j.s = s; j.sd = sd; j.t = tt; j.td = td; j.cov = 30; j.cn = 1; g.juncs.push_back(j); }
                else if (found->cn < 2) found->cn = 2;
                last = i;
            }
        }
        if (segs.size() - last >= 2) { std::vector<int> sub(segs.begin() + last, segs.end()); std::sort(sub.begin(), sub.end()); res.push_back(sub); }
    }
    std::sort(res.begin(), res.end());
    res.erase(std::unique(res.begin(), res.end()), res.end());
    return res;
}

// junctions that indelBFB / getIndelBias look at: inside the partition, not a fold-back, not a reference adjacency
static bool normal_adjacency(const Junc &j) { return j.sd == j.td && ((j.sd == '+' && j.t - j.s == 1) || (j.sd == '-' && j.s - j.t == 1)); }

// getIndelBias (LGM:3699-3744)
static void get_indel_bias(LhGraph &g, int s_id, int e_id)
{
    std::vector<Junc> sv;
    for (const Junc &j : g.juncs) {
        if (j.s < s_id || j.s > e_id || j.t < s_id || j.t > e_id) continue;
        if (j.sd != j.td || normal_adjacency(j)) continue;
        sv.push_back(j);
    }
    while (!sv.empty()) {
        std::vector<int> group;
        for (size_t i = 0; i < sv.size(); ++i) {
            int s = sv[i].sd == '-' ? -sv[i].s : sv[i].s, t = sv[i].td == '-' ? -sv[i].t : sv[i].t;
            if (group.empty()) { group.push_back(s); group.push_back(t); }
            else if (t == group.front()) group.insert(group.begin(), s);
            else if (s == -group.front()) group.insert(group.begin(), -t);
            else if (group.back() == s) group.push_back(t);
            else if (group.back() == -t) group.push_back(-s);
            else continue;
            sv.erase(sv.begin() + i); --i;
        }
        if (group.size() == 2) {
            if (group[0] < group[1]) { for (int k = group[0] + 1; k < group[1]; ++k) if (Seg *s = g.seg(std::abs(k))) s->cn += 1; }
            else { for (int k = group[1]; k <= group[0]; ++k) if (Seg *s = g.seg(std::abs(k))) s->cn -= 1; }
        } else {
            for (size_t k = 1; k + 1 < group.size(); ++k) if (Seg *s = g.seg(std::abs(group[k]))) s->cn -= 1;
        }
    }
}

typedef std::vector<int> Path;   // signed ids: +id '+', -id '-'

static Path::iterator find_from(Path &p, Path::iterator from, int v) { return from > p.end() ? p.end() : std::find(from, p.end(), v); }

// indelBFB (LGM:3746-3837): splice deletions / duplications / inversions / insertions back into the path
static bool indel_bfb(LhGraph &g, Path &path, int s_id, int e_id)
{
    std::vector<Junc> sv;
    for (const Junc &j : g.juncs) {
        if (j.s < s_id || j.s > e_id || j.t < s_id || j.t > e_id) continue;
        if (j.sd != j.td && std::abs(j.s - j.t) <= 2) continue;
        if (normal_adjacency(j)) continue;
        sv.push_back(j);
    }
    if (sv.empty()) return false;
    while (!sv.empty()) {
        Path group;
        for (size_t i = 0; i < sv.size(); ++i) {
            int as = sv[i].sd == '-' ? -sv[i].s : sv[i].s, at = sv[i].td == '-' ? -sv[i].t : sv[i].t;
            int bs = -at, bt = -as;                                  // complement edge (Junction.cpp:26-40)
            if (group.empty()) { group.push_back(as); group.push_back(at); }
            else if (at == group.front()) group.insert(group.begin(), as);
            else if (bt == group.front()) group.insert(group.begin(), bs);
            else if (group.back() == as) group.push_back(at);
            else if (group.back() == bs) group.push_back(bt);
            else continue;
            sv.erase(sv.begin() + i); --i;
        }
        auto flip = [&]() { std::reverse(group.begin(), group.end()); for (int &v : group) v = -v; };
        if (group.size() == 2) {
            const bool same_dir = (group[0] > 0) == (group[1] > 0);
            const bool deletion = same_dir && ((group[0] > 0 && std::abs(group[0]) < std::abs(group[1])) || (group[0] < 0 && std::abs(group[0]) > std::abs(group[1])));
            if (same_dir && !deletion) {                                  // duplication
                auto p1 = std::find(path.begin(), path.end(), group[0]);
                auto p2 = std::find(path.begin(), p1, group[1]);
                if (p1 == path.end() || p2 == p1) { flip(); p1 = std::find(path.begin(), path.end(), group[0]); p2 = std::find(path.begin(), p1, group[1]); }
                if (p1 == path.end() || p2 == p1) continue;
                Path copy(p2, p1 + 1);
                path.insert(p1 + 1, copy.begin(), copy.end());
            } else {                                                      // deletion or inversion
                const long limit = deletion ? 3 : 5;
                auto p1 = std::find(path.begin(), path.end(), group[0]);
                auto p2 = find_from(path, p1 + (p1 == path.end() ? 0 : 1), group[1]);
                if (p1 == path.end() || p2 == path.end()) { flip(); p1 = std::find(path.begin(), path.end(), group[0]); p2 = find_from(path, p1 + (p1 == path.end() ? 0 : 1), group[1]); }
                if (p1 == path.end() || p2 == path.end() || p2 - p1 > limit) continue;
                path.erase(p1 + 1, p2);
            }
        } else {                                                          // insertion
            auto p1 = std::find(path.begin(), path.end(), group.front());
            auto p2 = find_from(path, p1 + (p1 == path.end() ? 0 : 1), group.back());
            if (p1 == path.end() || p2 == path.end()) { flip(); p1 = std::find(path.begin(), path.end(), group.front()); p2 = find_from(path, p1 + (p1 == path.end() ? 0 : 1), group.back()); }
            if (p1 == path.end() || p2 == path.end()) continue;
            size_t at = p1 - path.begin();
            path.erase(p1 + 1, p2);
            path.insert(path.begin() + at + 1, group.begin() + 1, group.end() - 1);
        }
    }
    return true;
}

// printBFB (LGM:3411-3429) for a single-chromosome path
static std::string format_path(const Path &p)
{
    std::string out;
    for (size_t i = 0; i < p.size(); ++i) {
        out += std::to_string(std::abs(p[i])) + (p[i] > 0 ? "+" : "-");
        if (i + 1 < p.size() && (p[i] > 0) != (p[i + 1] > 0)) out += "|";
    }
    return out;
}

struct Partition {
    int s_id = 0, e_id = 0, bias = 1;
    bool trivial = false;
    std::vector<int32_t> c2, j2, tgs, inv_src, inv_tgt;
    int32_t loop_ub = 0;
    int unit_shift = 0;                 // copy numbers in units of 1/(2 << unit_shift) copies
    std::vector<int32_t> element_cn;
    LhGraph *graph = nullptr;
};

// getJuncCN (LGM:3989-4050) + bias (localhap.cpp:141-146) + getIndelBias + component selection (localhap.cpp:158-162)
static bool flatten(LhGraph &g, const std::vector<std::vector<int>> &components, std::vector<Partition> &out, std::string &why)
{
    int max_id = 0;
    for (const Seg &s : g.segs) max_id = std::max(max_id, s.id);
    for (size_t n = 0; n < g.sources.size(); ++n) {
        Partition p; p.graph = &g;
        p.s_id = g.sources[n]; p.e_id = g.sinks[n];
        const int len = p.e_id - p.s_id + 1;
        std::vector<double> fbi(max_id + 2, 0.0);
        p.inv_src.assign(max_id + 2, 0); p.inv_tgt.assign(max_id + 2, 0);
        std::vector<const Junc *> inv;
        for (const Junc &j : g.juncs) {
            if (j.s < p.s_id || j.s > p.e_id || j.t < p.s_id || j.t > p.e_id) continue;
            double cn = j.cn; if (cn > 0.5 && cn < 1) cn = 1;
            if (j.sd != j.td && std::abs(j.s - j.t) <= 2) {
                inv.push_back(&j);
                if (!p.inv_src[j.s]) { p.inv_src[j.s] = j.s; p.inv_tgt[j.s] = j.t; fbi[j.s] += cn; }
                else if (!p.inv_src[j.t]) { p.inv_src[j.t] = j.s; p.inv_tgt[j.t] = j.t; fbi[j.t] += cn; }
            }
        }
        for (const Junc *j : inv) {
            if (!p.inv_src[j->s]) { p.inv_src[j->s] = j->s; p.inv_tgt[j->s] = j->t; }
            if (!p.inv_src[j->t]) { p.inv_src[j->t] = j->s; p.inv_tgt[j->t] = j->t; }
        }
        for (int i = p.s_id; i <= p.e_id; ++i)
            if (fbi[i] > 0 && p.inv_src[i] != p.inv_tgt[i]) p.bias += (int)fbi[i] % 2;
        get_indel_bias(g, p.s_id, p.e_id);
        double fsum = 0;
        for (int i = 0; i <= p.e_id; ++i) fsum += fbi[i];
        std::vector<std::vector<int>> valid;
        for (const auto &c : components) { Seg *s = g.seg(c[0]); if (s && s->chr_id == (int)n) valid.push_back(c); }
        p.trivial = std::fabs(fsum) < 1e-6 && valid.empty();
        if (!p.trivial) {
            double total = 0;
            for (const Seg &s : g.segs) total += s.cn;
            p.loop_ub = (int32_t)std::floor(total);
            // The reference's model takes any real copy number (Weight.hpp:11, Graph.cpp:369-405).  The device searches
            // in integers: copy numbers in units of 1/(2U) copies, U = 2^unit_shift.  Take the coarsest grid the
            // partition's values lie on; values on no grid up to 1/128 copy (depth-derived decimals) are rounded to it.
            std::vector<double> cs, js;
            for (int i = p.s_id; i <= p.e_id; ++i) {
                Seg *s = g.seg(i);
                if (!s) { why = "missing segment " + std::to_string(i); return false; }
                if (s->cn < 0) { why = "negative copy number (segment " + std::to_string(i) + ")"; return false; }
                cs.push_back(s->cn); js.push_back(fbi[i]);
            }
            int ush = 0;
            for (; ush < 6; ++ush) {
                const double u = (double)(2 << ush);
                bool ok = true;
                for (size_t k = 0; k < cs.size() && ok; ++k)
                    ok = std::fabs(u * cs[k] - std::round(u * cs[k])) <= 1e-9 * u && std::fabs(u * js[k] - std::round(u * js[k])) <= 1e-9 * u;
                if (ok) break;
            }
            p.unit_shift = ush;
            const double u = (double)(2 << ush);
            for (size_t k = 0; k < cs.size(); ++k) { p.c2.push_back((int32_t)std::llround(u * cs[k])); p.j2.push_back((int32_t)std::llround(u * js[k])); }
            for (const auto &c : valid) { p.tgs.push_back(std::min(c.front(), c.back())); p.tgs.push_back(std::max(c.front(), c.back())); }
            p.element_cn.assign((size_t)len * (len + 1), 0);
        }
        out.push_back(p);
    }
    return true;
}

int main(int argc, char **argv)
{
    std::map<std::string, std::string> opt;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.rfind("--", 0) != 0) continue;
        a = a.substr(2);
        if (a == "reversed" || a == "all" || a == "help") { opt[a] = "true"; continue; }
        size_t eq = a.find('=');
        if (eq != std::string::npos) opt[a.substr(0, eq)] = a.substr(eq + 1);
        else if (i + 1 < argc) opt[a] = argv[++i];
    }
    if (opt.count("help") || !opt.count("op")) {
        std::cout << "usage: ambigram_bfb --op bfb --in_lh X.lh [--juncdb F --junc_info true] [--reversed] | --in_list LIST" << std::endl;
        return opt.count("help") ? 0 : 1;
    }
    std::cout << opt["op"] << std::endl;
    if (opt["op"] != "bfb") { std::cerr << "only --op bfb is supported" << std::endl; return 1; }
    const bool junc_info = opt["junc_info"] == "true" || opt["junc_info"] == "1";
    const bool reversed = opt.count("reversed") > 0;
    const bool print_all = opt.count("all") > 0;
    auto t0 = std::chrono::steady_clock::now();

    std::vector<std::string> files;
    if (opt.count("in_list")) { std::ifstream in(opt["in_list"]); std::string l; while (std::getline(in, l)) if (!l.empty()) files.push_back(l); }
    else if (opt.count("in_lh")) files.push_back(opt["in_lh"]);
    else { std::cerr << "--in_lh or --in_list is required" << std::endl; return 1; }

    std::vector<LhGraph> graphs(files.size());
    std::vector<std::vector<Partition>> parts(files.size());
    for (size_t f = 0; f < files.size(); ++f) {
        std::cout << "Reading graph..." << std::endl;
        if (!parse_lh(files[f], graphs[f])) return 1;
        std::vector<std::vector<int>> comps = read_components(graphs[f], files.size() == 1 ? opt["juncdb"] : std::string());
        std::string why;
        if (!flatten(graphs[f], comps, parts[f], why)) { std::cerr << files[f] << ": " << why << std::endl; return 1; }
    }

    ambi_ctx *ctx = nullptr;
    ambi_config cfg; memset(&cfg, 0, sizeof cfg);
    if (const char *d = getenv("AMBIGRAM_DEVICE")) cfg.device = atoi(d);
    // a single .lh needs little frontier memory (an overflowing level falls back to the chunked pass);
    // a list gets the library default (a fifth of the free HBM)
    if (files.size() == 1) cfg.node_capacity = 2 << 20;
    if (const char *c = getenv("AMBIGRAM_NODE_CAPACITY")) cfg.node_capacity = atoll(c);
    if (ambi_init(&ctx, &cfg) != AMBI_OK) { std::cerr << "ambi_init: " << ambi_last_error(nullptr) << std::endl; return 1; }

    // ---- Stage A for every non-trivial partition in one batch
    std::vector<Partition *> work;
    for (auto &pv : parts) for (Partition &p : pv) if (!p.trivial) work.push_back(&p);
    std::vector<ambi_bfb_problem> probs(work.size());
    std::vector<ambi_bfb_solution> sols(work.size());
    for (size_t i = 0; i < work.size(); ++i) {
        Partition &p = *work[i];
        memset(&probs[i], 0, sizeof probs[i]); memset(&sols[i], 0, sizeof sols[i]);
        probs[i].n = p.e_id - p.s_id + 1; probs[i].start_id = p.s_id; probs[i].seg_cn2 = p.c2.data(); probs[i].fbi_cn2 = p.j2.data();
        probs[i].loop_ub = p.loop_ub; probs[i].bias = p.bias; probs[i].n_tgs = (int32_t)(p.tgs.size() / 2);
        probs[i].tgs_pairs = p.tgs.empty() ? nullptr : p.tgs.data(); probs[i].flags = junc_info ? AMBI_FLAG_JUNC_INFO : 0u;
        probs[i].unit_shift = (uint32_t)p.unit_shift;
        sols[i].element_cn = p.element_cn.data();
    }
    if (!work.empty() && ambi_bfb_solve_batch(ctx, (int32_t)work.size(), probs.data(), sols.data()) != AMBI_OK) {
        std::cerr << "ambi_bfb_solve_batch: " << ambi_last_error(ctx) << std::endl; return 1;
    }
    // ---- Stage B
    std::vector<ambi_dag> dags(work.size());
    std::vector<int32_t> rev(work.size(), reversed ? 1 : 0);
    std::vector<ambi_order_result> res(work.size());
    for (size_t i = 0; i < work.size(); ++i)
        if (ambi_bfb_construct_dag(probs[i].n, probs[i].start_id, work[i]->element_cn.data(), &dags[i]) != AMBI_OK) { std::cerr << "construct_dag failed" << std::endl; return 1; }
    if (!work.empty() && ambi_bfb_first_valid_order_batch(ctx, (int32_t)work.size(), dags.data(), rev.data(), res.data()) != AMBI_OK) {
        std::cerr << "ambi_bfb_first_valid_order_batch: " << ambi_last_error(ctx) << std::endl; return 1;
    }
    // ---- host post-processing and printing, file by file, partition by partition
    size_t wi = 0;
    double ilp_error = 0;
    for (size_t f = 0; f < files.size(); ++f) {
        int path_len = 0;
        for (Partition &p : parts[f]) {
            Path path;
            if (p.trivial) { for (int i = p.s_id; i <= p.e_id; ++i) path.push_back(i); std::cout << format_path(path) << std::endl; path_len += (int)path.size(); continue; }
            ambi_order_result &r = res[wi];
            ilp_error += sols[wi].objective2 / (double)(2 << probs[wi].unit_shift);      // what cbc prints as "objective value"
            ++wi;
            if (print_all && r.found) {
                // --all: every valid order in the reference's print order (LGM:3672-3695); `path` keeps the first
                // (enumeration, imperfectFBI and expansion all on the device: ambi_bfb_all_valid_paths)
                const int kmax = 4096, pstride = 8192;
                std::vector<int32_t> pv((size_t)kmax * pstride), pl(kmax);
                int32_t n0 = 0, n1 = 0;
                if (ambi_bfb_all_valid_paths(ctx, &dags[wi - 1], reversed ? 1 : 0, p.inv_src.data(), p.inv_tgt.data(), (int32_t)p.inv_src.size() - 2,
                                             kmax, nullptr, pv.data(), pstride, pl.data(), &n0, &n1, nullptr) != AMBI_OK) {
                    std::cerr << "ambi_bfb_all_valid_paths: " << ambi_last_error(ctx) << std::endl; return 1;
                }
                for (int q = 0; q < n0 + n1; ++q) {
                    Path one(pv.begin() + (size_t)q * pstride, pv.begin() + (size_t)q * pstride + pl[q]);
                    if (q == 0) path = one;
                    std::cout << format_path(one) << std::endl;
                }
            } else if (r.found) {
                ambi_bfb_imperfect_fbi(r.bkp, r.bkp_len, p.inv_src.data(), p.inv_tgt.data(), (int32_t)p.inv_src.size() - 2);
                std::vector<int32_t> buf(1 << 16);
                int len = ambi_bfb_expand_path(r.bkp, r.bkp_len, buf.data(), (int32_t)buf.size());
                path.assign(buf.begin(), buf.begin() + std::min<int>(len, (int)buf.size()));
                std::cout << format_path(path) << std::endl;
            }
            if (indel_bfb(*p.graph, path, p.s_id, p.e_id)) {
                std::cout << "BFB path with insertion, deletion, or duplication:\n" << format_path(path) << std::endl;
            }
            path_len += (int)path.size();
        }
        // time.csv row (localhap.cpp:382-388)
        LhGraph &g = graphs[f];
        int ninv = 0; double cn_sum = 0, max_cn = 0;
        for (Partition &p : parts[f]) { std::set<int> ids; for (size_t i = 0; i < p.inv_src.size(); ++i) if (p.inv_src[i]) ids.insert((int)i); ninv += (int)ids.size(); }
        for (const Seg &s : g.segs) { cn_sum += (int)s.cn; max_cn = std::max(max_cn, std::floor(s.cn)); }
        if (!getenv("AMBIGRAM_NO_TIME_CSV")) {
            std::ofstream tf("time.csv", std::ios_base::app);
            double sec = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() / 1e6;
            tf << files[f].substr(0, files[f].find(".")) << "," << g.segs.size() << "," << ninv << "," << (int)g.juncs.size() - ninv << ","
               << (int)cn_sum << "," << path_len << "," << (int)max_cn << "," << sec << "\n";
        }
    }
    (void)ilp_error;
    ambi_destroy(ctx);
    return 0;
}
