// PoseFiles.cpp -- see PoseFiles.h.  Pure host code: streams in, streams out.
#include "PoseFiles.h"

#include <cstring>
#include <fstream>
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

bool read_batch_config(std::istream& is) {                                                          // src/main.cpp:132-137
    int start = 0, size = 0, bootstrap = 0;
    is >> start >> size >> bootstrap;
    if (is.fail()) return false;
    util::BATCH_START_ID = start;
    util::BATCH_SIZE = size;
    util::FLAG_IS_BOOTSTRAP = bootstrap != 0;
    return true;
}

int configure_from_args(int argc, const char* const* argv, std::string* message) {                  // src/main.cpp:80-101
    if (argc == 2) { if (message) *message = "Either Config. file or loop closure flag missing! Exiting..."; return -1; }
    if (argc != 3) return 0;
    if (!std::strcmp(argv[1], "LC")) util::FLAG_ALTERNATE_GN_RA = true;
    std::ifstream f(argv[2]);
    if (!f.is_open()) { if (message) *message = "Unable to open Config. file! Exiting..."; return -1; }
    if (!util::FLAG_ALTERNATE_GN_RA) return 0;                                                      // the file is only read in batch mode
    if (!read_batch_config(f)) { if (message) *message = "Config. file does not hold BATCH_START_ID BATCH_SIZE FLAG_IS_BOOTSTRAP"; return -1; }
    return 1;
}

}  // namespace ellc_host
