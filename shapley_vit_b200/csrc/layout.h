// Plan layout table (see layout.cu / shapley_vit_b200/layout.py).
#pragma once
#include <vector>

#include "common.cuh"

namespace svit {

struct Layout {
  svit_vit_cfg cfg;
  std::vector<svit_segment> segs;
  int64_t vec_size = 0, mat_size = 0;
  // element offset of (kind, layer) inside its region, or -1
  int64_t find(int kind, int layer = -1) const;
};

int validate_cfg(const svit_vit_cfg* c);
int build_layout(const svit_vit_cfg* c, Layout* L);

}  // namespace svit
