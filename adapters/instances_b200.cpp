// instances_b200.cpp -- the registration unit a maintainer adds next to R/instances.cpp (pattern: :21-23 the
// linker-friendly macro, :28-85 one BOSS_REGISTER_CLASS per concrete class): after
// srrg2_slam_interfaces_b200_registerTypes() the B200 classes can be named in a BOSS configuration wherever the
// stock finder / aligner / solver classes are named today -- the configuration selects them by class name, nothing
// else in srrg2_laser_slam_2d / srrg2_proslam changes.  Compiled here against the Appendix-A stub headers (whose
// BOSS_REGISTER_CLASS is a no-op): the BOSS text round trip itself needs srrg2_core's serializer and can
// only be exercised on a machine that has it.
#include "correspondence_finder_b200.h"
#include "multi_aligner_b200.h"
#include "solver_b200.h"

#define BOSS_REGISTER_CLASS_LINKER_FRIENDLY(CLASS_) \
  CLASS_ dummy_##CLASS_;                            \
  BOSS_REGISTER_CLASS(CLASS_)

namespace srrg2_slam_interfaces {

void srrg2_slam_interfaces_b200_registerTypes() {
  BOSS_REGISTER_CLASS_LINKER_FRIENDLY(CorrespondenceFinderB2002D);
  BOSS_REGISTER_CLASS_LINKER_FRIENDLY(CorrespondenceFinderB2003D);
  BOSS_REGISTER_CLASS_LINKER_FRIENDLY(MultiAligner2DB200);
  BOSS_REGISTER_CLASS_LINKER_FRIENDLY(MultiAligner3DQRB200);
  BOSS_REGISTER_CLASS_LINKER_FRIENDLY(PoseGraphSolver2DB200);
  BOSS_REGISTER_CLASS_LINKER_FRIENDLY(PoseGraphSolver3DB200);
}

}  // namespace srrg2_slam_interfaces
