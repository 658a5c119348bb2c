#!/usr/bin/env python3
"""oracle/gen_hooks.py -- TEST INFRASTRUCTURE.
Derives the mangled names of the reference functions that arriba.cpp:79 `main` calls (from the
compiled, unmodified arriba.o) and writes
  _ref/hook_syms.h   #define MANGLED_<fn> "<mangled>"   (used by dump_hooks.cpp asm labels)
  _ref/wrap.flags    -Wl,--wrap=<mangled> ...           (GNU ld symbol interposition)
so that oracle/dump_hooks.cpp can observe the containers between the reference's own stages
without touching a single reference source line."""
import subprocess, sys, re
obj, out_h, out_flags = sys.argv[1:4]
HOOKED = """mark_multimappers detect_strandedness estimate_fragment_length read_chimeric_alignments
filter_duplicates filter_uninteresting_contigs filter_viral_contigs filter_top_expressed_viral_contigs
filter_low_coverage_viral_contigs filter_proximal_read_through filter_inconsistently_clipped_mates
filter_homopolymer filter_small_insert_size filter_long_gap filter_same_gene filter_hairpin
filter_mismatches filter_low_entropy find_fusions merge_adjacent_fusions filter_multimappers
estimate_expected_fusions filter_non_coding_neighbors filter_intragenic_both_exonic filter_min_support
filter_relative_support recover_internal_tandem_duplication filter_both_intronic filter_in_vitro
recover_both_spliced select_most_supported_breakpoints filter_marginal_read_through recover_many_spliced
filter_short_anchor filter_end_to_end_fusions filter_no_coverage make_kmer_index filter_homologs
filter_mismappers recover_isoforms assign_confidence write_fusions_to_file""".split()
syms = subprocess.check_output(["nm", "-u", obj], text=True).split()
found = {}
for s in syms:
    m = re.match(r"_Z(\d+)([a-z_]+)", s)
    if m and len(m.group(2)) >= int(m.group(1)):
        name = m.group(2)[:int(m.group(1))]
        if name in HOOKED:
            assert name not in found, "overloaded: " + name
            found[name] = s
missing = [h for h in HOOKED if h not in found]
assert not missing, "not referenced by main: %s" % missing
with open(out_h, "w") as f:
    for k in HOOKED:
        f.write('#define MANGLED_%s "%s"\n' % (k, found[k]))
with open(out_flags, "w") as f:
    f.write(" ".join("-Wl,--wrap=" + found[k] for k in HOOKED) + "\n")
