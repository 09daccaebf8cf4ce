// COIN-OR header shim (TEST INFRASTRUCTURE): the BFB path never solves in-process
// (src/LocalGenomicMap.cpp:751 is in the dead balancerILP block); only the type must exist.
#pragma once
#include "OsiClpSolverInterface.hpp"
class CbcModel {
public:
    explicit CbcModel(OsiClpSolverInterface &) {}
};
