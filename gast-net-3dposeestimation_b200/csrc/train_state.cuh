// Saved-for-backward state of one training forward (see train.cuh).
#pragma once
#include <string>
#include <vector>
#include "gast_common.cuh"

using gast::RowMap;

struct BnSave {
  const float* Z = nullptr; int ldz = 0; long long M = 0; int N = 0;
  float* mean = nullptr; float* invstd = nullptr;
  const float* gamma = nullptr; const float* beta = nullptr;
  int relu = 0; const unsigned char* keep = nullptr; float kscale = 1.f;
  std::string prefix;
};

struct BlockSave {
  int C = 0; long long F = 0;
  const float* X = nullptr;
  float *Wst = nullptr, *H = nullptr, *coefA[2] = {nullptr, nullptr}, *S = nullptr, *XY = nullptr, *Zlc = nullptr, *L = nullptr;
  float *G = nullptr, *AB = nullptr, *Y = nullptr, *Zgc = nullptr, *Gl = nullptr, *Zbc = nullptr, *Out = nullptr;
  BnSave bn1, bn2, bnlc, bngc, bnbc;
  std::string P;
};

struct StageSave {
  int Cw = 0, taps = 0; long long Fin = 0, Fout = 0; int Tin = 0, Tout = 0;
  const float* X = nullptr; float *Wt = nullptr, *Z1 = nullptr, *Hh = nullptr, *Z2 = nullptr, *Out = nullptr;
  RowMap tapmap, resmap; int dil = 1;
  BnSave bnA, bnB;
  int idx = 0;
};

struct TrainState {
  bool valid = false;
  int B = 0, T = 0, T0 = 0, s0 = 1;   // s0: stride of the expand conv (filter width when strided, else 1)
  const float* x = nullptr;
  float *xbn = nullptr, *Acol = nullptr, *We8 = nullptr, *Z0 = nullptr, *act0 = nullptr;
  BnSave bnin, bnex;
  std::vector<BlockSave> blocks;
  std::vector<StageSave> stages;
  const float* last = nullptr;   // input of shrink
  long long Flast = 0;
  float drop_p = 0.f; unsigned long long seed = 0; unsigned long long drop_ctr = 0;
  size_t arena_off = 0;          // workspace used by the forward (backward temporaries go after it)
};

