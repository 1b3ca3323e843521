// PoseFiles.h -- the text files the reference's callers write from the tracker's results and that the unmodified MATLAB
// rotation averaging reads.  Same columns, same default-ostream number formatting (6 significant digits),
// same id offset (frameId + util::BATCH_START_ID - 1).
//
//   poses_orig.txt         src/main.cpp:368-375   frameId kfId poseWrtWorld[6] rescaleFactor depthMapOccupancy
//   matchframes.txt        src/main.cpp:378-384   frameId kfId poseWrtOrigin[6] rescaleFactor seeds 0 0 0
//   matchframes_globalopt  src/GlobalOptimize.cpp:574-582   ... seeds matchValue rms_error relative_view_angle
//   initial poses (input)  src/main.cpp:207-210   frame_no pose[6]
#pragma once

#include <iosfwd>

#include "Frame.h"

namespace ellc_host {

void write_orig_pose(std::ostream& os, const frame* f, const frame* keyframe, float seeds_num);
void write_match_pose(std::ostream& os, const frame* f, const frame* keyframe, float seeds_num);                 // sequential pair: "0 0 0" tail
void write_match_pose(std::ostream& os, const frame* f, const frame* keyframe, float seeds_num, float matchValue,
                      float rms_error, float relative_view_angle);                                               // loop-closure pair
bool read_initial_pose(std::istream& is, int& frame_no, float pose[6]);

}  // namespace ellc_host
