#!/bin/bash
# Mutated JPEG streams through the header parser and the host model of the decode kernels (the entropy-decoding core the
# kernels compile), built with AddressSanitizer + UBSan.  CPU only.  usage: tools/fuzz_jpeg_asan.sh [streams-per-case]
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
N=${1:-40}
W=$(mktemp -d)
python - "$ROOT" "$W" "$N" <<'PY'
import sys
root, out, per = sys.argv[1], sys.argv[2], int(sys.argv[3])
sys.path.insert(0, root); sys.path.insert(0, root + "/tests")
import numpy as np
from jpeg_cases import make_cases
sc, streams = make_cases()
rng = np.random.default_rng(7); n = 0
for name in streams:
    base = bytearray(streams[name])
    for t in range(per):
        bad = bytearray(base); k = t % 3
        if k == 0:
            for i in rng.integers(700, len(bad) - 2, size=int(rng.integers(1, 50))): bad[i] = int(rng.integers(0, 256))
        elif k == 1:
            del bad[int(rng.integers(100, len(bad))):]
        else:
            for i in rng.integers(2, 650, size=int(rng.integers(1, 6))): bad[i] = int(rng.integers(0, 256))
        open(f"{out}/{n:04d}.jpg", "wb").write(bytes(bad)); n += 1
print(n, "streams")
PY
cat > "$W/h.cc" <<'CC'
#include "jpeg.h"
#include <cstdio>
#include <cstring>
#include <vector>
using namespace b200tag;
int main(int argc, char **argv) {
  int ok = 0, rej = 0;
  for (int i = 1; i < argc; i++) {
    FILE *f = fopen(argv[i], "rb");
    std::vector<uint8_t> d;
    uint8_t b[65536];
    size_t n;
    while ((n = fread(b, 1, sizeof b, f)) > 0) d.insert(d.end(), b, b + n);
    fclose(f);
    uint8_t *p = new uint8_t[d.size()];  // exact-size heap copy: over-reads are caught
    memcpy(p, d.data(), d.size());
    std::vector<uint8_t> out(648 * 488);
    int rounds = 0;
    (jpeg_model_decode(p, d.size(), out.data(), out.size(), &rounds) == 0 ? ok : rej)++;
    delete[] p;
  }
  printf("decoded %d, rejected %d, no sanitizer report\n", ok, rej);
}
CC
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -ffp-contract=off -I"$ROOT/ros_vision_b200/csrc" \
  "$W/h.cc" "$ROOT/ros_vision_b200/csrc/jpeg_host.cc" -o "$W/h"
"$W/h" "$W"/*.jpg
rm -rf "$W"
