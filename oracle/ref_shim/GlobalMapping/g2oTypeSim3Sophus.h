// oracle/ref_shim/GlobalMapping/g2oTypeSim3Sophus.h -- TEST INFRASTRUCTURE.
// DataStructures/FramePoseStruct.h:23,59 needs the NAME VertexSim3 for a pointer member that stays null outside the
// pose-graph thread; g2o is absent and out of scope.
#pragma once
namespace lsd_slam { class VertexSim3; }
