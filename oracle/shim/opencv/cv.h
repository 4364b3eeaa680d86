// oracle/shim: forwards to the oracle cv shim (TEST INFRASTRUCTURE; lets the reference compile verbatim without OpenCV)
#include "cv_shim_all.hpp"
