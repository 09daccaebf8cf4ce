// COIN-OR header shim (TEST INFRASTRUCTURE, oracle/Makefile target `ref`).
//
// The reference's BFB path uses COIN-OR only as an LP/MPS *file writer*
// (src/LocalGenomicMap.cpp:4400, 4420, 4742-4750); the solve itself is an external
// `cbc` process (localhap.cpp:179-181).  COIN-OR is not installed in this image, so
// this shim provides the handful of classes the reference touches.  writeLp() dumps
// the model the reference built -- rows, bounds, objective, integrality -- in a plain
// text format ("AMBI-LP 1") that oracle/cbc_standin.py solves with HiGHS and that
// the tests compare row for row with oracle/ambi_oracle.c.
#pragma once
#include <vector>
#include <string>
#include <cstdio>
#include <limits>

class CoinPackedVector {
public:
    std::vector<int> idx;
    std::vector<double> val;
    void insert(int i, double v) { idx.push_back(i); val.push_back(v); }
};

class CoinPackedMatrix {
public:
    std::vector<CoinPackedVector> rows;
    CoinPackedMatrix(bool, int, int) {}
    void appendRow(const CoinPackedVector &r) { rows.push_back(r); }
};

class OsiClpSolverInterface {
    const CoinPackedMatrix *m_ = nullptr;
    std::vector<double> clb_, cub_, obj_, rlb_, rub_;
    std::vector<char> integer_;
    int ncols_ = 0;

public:
    double getInfinity() const { return std::numeric_limits<double>::infinity(); }
    void loadProblem(const CoinPackedMatrix &m, const double *collb, const double *colub, const double *obj,
                     const double *rowlb, const double *rowub) {
        m_ = &m;
        int maxc = -1;
        for (const auto &r : m.rows) for (int i : r.idx) if (i > maxc) maxc = i;
        ncols_ = maxc + 1;
        clb_.assign(collb, collb + ncols_); cub_.assign(colub, colub + ncols_); obj_.assign(obj, obj + ncols_);
        rlb_.assign(rowlb, rowlb + m.rows.size()); rub_.assign(rowub, rowub + m.rows.size());
        integer_.assign(ncols_, 0);
    }
    void setInteger(int i) { if (i >= 0 && i < ncols_) integer_[i] = 1; }
    bool isInteger(int i) const { return i >= 0 && i < ncols_ && integer_[i]; }
    void writeMps(const char *) {}
    void writeLp(const char *prefix) {
        std::string fn = std::string(prefix) + ".lp";
        FILE *f = fopen(fn.c_str(), "w");
        if (!f) return;
        fprintf(f, "AMBI-LP 1\n%zu %d\n", m_->rows.size(), ncols_);
        for (size_t r = 0; r < m_->rows.size(); ++r) {
            fprintf(f, "R %.17g %.17g %zu", rlb_[r], rub_[r], m_->rows[r].idx.size());
            for (size_t k = 0; k < m_->rows[r].idx.size(); ++k) fprintf(f, " %d:%.17g", m_->rows[r].idx[k], m_->rows[r].val[k]);
            fprintf(f, "\n");
        }
        for (int c = 0; c < ncols_; ++c) fprintf(f, "C %.17g %.17g %.17g %d\n", clb_[c], cub_[c], obj_[c], (int)integer_[c]);
        fclose(f);
    }
    const double *getColSolution() const { return nullptr; }
};
