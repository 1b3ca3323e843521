// PoseFiles.h -- the text files the reference's callers write from the tracker's results and that the unmodified MATLAB
// rotation averaging reads.  Same columns, same default-ostream number formatting (6 significant digits),
// same id offset (frameId + util::BATCH_START_ID - 1).
//
//   poses_orig.txt         src/main.cpp:368-375   frameId kfId poseWrtWorld[6] rescaleFactor depthMapOccupancy
//   matchframes.txt        src/main.cpp:378-384   frameId kfId poseWrtOrigin[6] rescaleFactor seeds 0 0 0
//   matchframes_globalopt  src/GlobalOptimize.cpp:574-582   ... seeds matchValue rms_error relative_view_angle
//   initial poses (input)  src/main.cpp:207-210   frame_no pose[6]
//   config.txt (input)     src/main.cpp:89-101, :132-137   BATCH_START_ID BATCH_SIZE FLAG_IS_BOOTSTRAP (written by bin/ELLC_LC.sh
//                          between the Gauss-Newton and rotation-averaging halves of a batch)
#pragma once

#include <iosfwd>
#include <string>

#include "Frame.h"

namespace ellc_host {

void write_orig_pose(std::ostream& os, const frame* f, const frame* keyframe, float seeds_num);
void write_match_pose(std::ostream& os, const frame* f, const frame* keyframe, float seeds_num);                 // sequential pair: "0 0 0" tail
void write_match_pose(std::ostream& os, const frame* f, const frame* keyframe, float seeds_num, float matchValue,
                      float rms_error, float relative_view_angle);                                               // loop-closure pair
bool read_initial_pose(std::istream& is, int& frame_no, float pose[6]);
// The batch parameters of the "LC" mode: `ELLC LC config.txt` sets util::FLAG_ALTERNATE_GN_RA and reads the three values with
// `my_file >> util::BATCH_START_ID >> util::BATCH_SIZE >> util::FLAG_IS_BOOTSTRAP` (src/main.cpp:132-137).  Returns false if the
// stream does not hold three integers; the util:: variables are only written on success.
bool read_batch_config(std::istream& is);
// main()'s argument handling (src/main.cpp:80-101): argc == 3 and argv[1] == "LC" selects the batch mode and opens argv[2].
// Returns 0 = sequential mode (no arguments), 1 = batch mode configured from the file, -1 = error (message as the reference prints).
int configure_from_args(int argc, const char* const* argv, std::string* message);

}  // namespace ellc_host
