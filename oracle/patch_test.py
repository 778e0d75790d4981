#!/usr/bin/env python3
"""Generate oracle/_ref/test_SimulationNBody_b200.cpp and oracle/_ref/test_CUDABodies_b200.cpp: the reference's OWN
murb-test bodies with nothing but the class under test re-targeted at gpu+b200.

  src/test/implem/test_SimulationNBody.cpp   the compile-time target switch (lines 36-43: CUDABodiesAllocator +
                                             SimulationNBodyCUDATileFullDevice, or the OpenMP fallback) becomes
                                             B200BodiesAllocator + SimulationNBodyB200; loop, sections, tolerances
                                             and the golden model (SimulationNBodyNaive) stay the reference's
  src/test/implem/test_CUDABodies.cpp        `CUDABodies<float> cudaBodies(n, scheme)` becomes `B200Bodies cudaBodies(...)`

Like patch_main.py the edit is applied to the files where they lie under /root/reference and written to oracle/_ref/
(git-ignored), so no reference source enters the repository.  Every edited line is printed.
"""
import re
import sys

ref_root, out_dir = sys.argv[1], sys.argv[2]


def fail(msg):
    sys.exit("patch_test.py: " + msg)


# ---- test_SimulationNBody.cpp: swap the #ifdef USE_CUDA ... #else ... #endif block that declares simuTest
path = ref_root + "/src/test/implem/test_SimulationNBody.cpp"
src = open(path).read().split("\n")
start = next((i for i, l in enumerate(src) if l.strip() == "#ifdef USE_CUDA" and "targetAllocator" in src[i + 1]), None)
if start is None:
    fail("target switch not found in " + path)
end = next(i for i in range(start, len(src)) if src[i].strip() == "#endif")
replaced = src[start:end + 1]
if not any("SimulationNBodyCUDATileFullDevice<float> simuTest" in l for l in replaced):
    fail("unexpected target switch in " + path)
new_block = ["    B200BodiesAllocator targetAllocator(n, scheme);              // gpu+b200 (was lines %d-%d)" % (start + 1, end + 1),
             "    SimulationNBodyB200 simuTest(targetAllocator, soft);"]
out = src[:start] + new_block + src[end + 1:]
inc = next(i for i, l in enumerate(out) if '#include "SimulationNBodyNaive.hpp"' in l)
out.insert(inc, '#include "SimulationNBodyB200.hpp" // gpu+b200')
open(out_dir + "/test_SimulationNBody_b200.cpp", "w").write("\n".join(out))
print("patch_test: test_SimulationNBody.cpp lines %d-%d -> B200BodiesAllocator + SimulationNBodyB200 (%d of %d lines unchanged)"
      % (start + 1, end + 1, len(src) - len(replaced), len(src)))

# ---- test_CUDABodies.cpp: the device container under test
path = ref_root + "/src/test/implem/test_CUDABodies.cpp"
text = open(path).read()
text, n = re.subn(r"CUDABodies<float>\s+cudaBodies\(", "B200Bodies cudaBodies(", text)
if n != 2:
    fail("expected 2 CUDABodies<float> declarations in %s, found %d" % (path, n))
text = text.replace('#include "core/CUDABodies.hpp"', '#include "core/CUDABodies.hpp"\n#include "SimulationNBodyB200.hpp" // gpu+b200', 1)
open(out_dir + "/test_CUDABodies_b200.cpp", "w").write(text)
print("patch_test: test_CUDABodies.cpp: %d declarations re-targeted at B200Bodies" % n)
