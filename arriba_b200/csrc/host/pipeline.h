// pipeline.h -- host-side orchestration of one arriba run: reference loading, BAM ingest, annotation, device stages
// (through the public C ABI only), event-level logic and output. Mirrors the call sequence of the reference's
// main (arriba.cpp:79-631).
#pragma once
#include <string>
#include <vector>
#include "refdata.h"
#include "ingest.h"
#include "../../../include/arriba_b200.h"

namespace arb { namespace host {

struct run_options { // options_t (options.hpp:25-69) restricted to what this implementation consumes
	std::string bam_file, gtf_file, assembly_file, output_file, discarded_output_file;
	std::string interesting_contigs, viral_contigs;
	arb_params params;
	int strandedness;          // 0 no, 1 yes, 2 reverse, 3 auto
	unsigned fragment_length;  // -F 200
	int threads;               // host threads for decode/annotation (the reference's -@ only affects BAM decoding)
	int device;
	run_options();
};

struct pipeline {
	run_options opt; int threads;
	refdata ref; fragment_table frags; coverage_windows coverage; ingest_stats istats;
	arb_ctx* ctx;
	int strandedness; i32 max_mate_gap; float read_length_mean, mate_gap_mean, mate_gap_stddev; bool fragment_length_ok;
	std::vector<u8> labels, early;
	std::string log; // the reference's progress lines (arriba.cpp:61-66 style, without time stamps)
	double t_reference, t_ingest, t_annotate, t_upload, t_read_filters, t_fragment_length, t_find_fusions;
	pipeline(): ctx(NULL), strandedness(0), max_mate_gap(0), read_length_mean(0), mate_gap_mean(0), mate_gap_stddev(0), fragment_length_ok(false) {}
	~pipeline();
	void load_reference();
	void ingest();
	void annotate();
	void upload();
	void read_filters();
	void fragment_length();
	void find_fusions();
	void run_all();
};

int detect_strandedness(pipeline& p);
void assign_strands(pipeline& p, int strandedness);
void annotate_fragments(pipeline& p);
bool estimate_fragment_length(pipeline& p, const u8* early, float& gap_mean, float& gap_stddev, float& read_length_mean);

}} // namespace
