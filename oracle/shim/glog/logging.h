// oracle/shim: stand-in for <glog/logging.h> (needlessly included by the reference's inc/ORBextractor.h:26).
// TEST INFRASTRUCTURE.
#ifndef ORACLE_SHIM_GLOG_H_
#define ORACLE_SHIM_GLOG_H_
#include <iostream>
namespace oracle_shim { struct NullLog { template <typename T> NullLog& operator<<(const T&) { return *this; } NullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } }; }
#define LOG(severity) ::oracle_shim::NullLog()
namespace google { inline void InitGoogleLogging(const char*) {} }
#endif
