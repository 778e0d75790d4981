#!/usr/bin/env python3
"""Generate oracle/_ref/main_b200.cpp = the reference's src/murb/main.cpp + the `gpu+b200` registration.

The reference selects implementations with an if/else-if chain on the --im string inside createImplem<T>()
(src/murb/main.cpp:205-270); there is no plugin table, so adding a tag means adding one #include and one branch.
This script applies exactly that edit to the file where it lies under /root/reference and writes the result to
oracle/_ref/ (git-ignored) so no reference source is copied into the repository.  INTEGRATION.md shows the same
edit as a diff for maintainers.  When MPI is not installed (this image) the `mpi` tag is dropped as well, because
SimulationNBodyMultiNode.hpp includes <mpi.h>.
"""
import sys

ref_main, out_path, have_mpi = sys.argv[1], sys.argv[2], sys.argv[3] == "1"
src = open(ref_main).read().split("\n")
out = []
i = 0
added_include = added_branch = added_help = False
while i < len(src):
    line = src[i]
    if not have_mpi and '#include "implem/SimulationNBodyMultiNode.hpp"' in line:
        i += 1
        continue
    if not have_mpi and 'ImplTag == "mpi"' in line:
        # drop `else if (ImplTag == "mpi") { ... }` (3 lines)
        while "}" not in src[i]:
            i += 1
        i += 1
        continue
    out.append(line)
    if '#include "implem/SimulationNBodyOpenMP.hpp"' in line and not added_include:
        out.append('#include "SimulationNBodyB200.hpp" // gpu+b200')
        added_include = True
    if '"gpu+leapfrog' in line and "docArgs" not in line and not added_help and "\\t" in line:
        out.append('                     "\\t\\t\\t - \\"gpu+b200\\n"')
        out.append('                     "\\t\\t\\t - \\"gpu+b200+leapfrog\\n"')
        added_help = True
    # the new branch goes right after the cpu+omp branch, ahead of the optional mpi / USE_CUDA branches
    if 'simu = new SimulationNBodyOpenMP<T>(allocator, Softening);' in line and not added_branch:
        out.append(src[i + 1])  # closing brace of the cpu+omp branch
        i += 1
        out += [
            '    else if (ImplTag == "gpu+b200" || ImplTag == "gpu+b200+leapfrog") {',
            '        B200BodiesAllocator b200Allocator(NBodies, BodiesScheme);',
            '        simu = new SimulationNBodyB200(b200Allocator, Softening, ImplTag == "gpu+b200+leapfrog");',
            '    }',
        ]
        added_branch = True
    i += 1
if not (added_include and added_branch):
    sys.exit("patch_main.py: anchors not found in " + ref_main)
open(out_path, "w").write("\n".join(out))
print("wrote", out_path, "(mpi kept)" if have_mpi else "(mpi tag dropped: no MPI in this image)")
