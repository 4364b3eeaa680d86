// include/ORBExtractor.h -- the reference's demo programs use a second, global-namespace copy of the class
// (reference inc/ORBExtractor.h:40-76, 5-argument operator()).  This header offers the same spelling on top
// of the single implementation in ORBextractor.h.
#ifndef ORBEXTRACTOR_GLOBAL_ALIAS_H
#define ORBEXTRACTOR_GLOBAL_ALIAS_H
#include "ORBextractor.h"
using ORB_SLAM3::ExtractorNode;
using ORB_SLAM3::ORBextractor;
#endif
