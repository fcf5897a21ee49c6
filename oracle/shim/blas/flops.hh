// oracle/shim: forwards to blas.hh (BLAS++ sub-header stand-in; test infrastructure only).
#pragma once
#include "../blas.hh"
