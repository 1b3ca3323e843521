// PoseFiles.cpp -- see PoseFiles.h.  Pure host code: streams in, streams out.
#include "PoseFiles.h"

#include <istream>
#include <ostream>

#include "ExternVariable.h"

namespace ellc_host {

static void ids(std::ostream& os, const frame* f, const frame* kf) {
    os << (f->frameId + util::BATCH_START_ID - 1) << " " << (kf->frameId + util::BATCH_START_ID - 1);
}

void write_orig_pose(std::ostream& os, const frame* f, const frame* kf, float seeds_num) {          // src/main.cpp:373
    ids(os, f, kf);
    for (int i = 0; i < 6; ++i) os << " " << f->poseWrtWorld[i];
    os << " " << kf->rescaleFactor << " " << seeds_num << "\n";
}

void write_match_pose(std::ostream& os, const frame* f, const frame* kf, float seeds_num) {         // src/main.cpp:382
    ids(os, f, kf);
    for (int i = 0; i < 6; ++i) os << " " << f->poseWrtOrigin[i];
    os << " " << kf->rescaleFactor << " " << seeds_num << " " << "0" << " " << "0" << " " << "0" << "\n";
}

void write_match_pose(std::ostream& os, const frame* f, const frame* kf, float seeds_num, float matchValue, float rms_error,
                      float relative_view_angle) {                                                  // src/GlobalOptimize.cpp:580
    ids(os, f, kf);
    for (int i = 0; i < 6; ++i) os << " " << f->poseWrtOrigin[i];
    os << " " << kf->rescaleFactor << " " << seeds_num << " " << matchValue << " " << rms_error << " " << relative_view_angle << "\n";
}

bool read_initial_pose(std::istream& is, int& frame_no, float pose[6]) {                            // src/main.cpp:210
    is >> frame_no >> pose[0] >> pose[1] >> pose[2] >> pose[3] >> pose[4] >> pose[5];
    return !is.fail();
}

}  // namespace ellc_host
