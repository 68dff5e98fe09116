// TEST INFRASTRUCTURE ONLY -- function-level driver around the UNMODIFIED reference
// split-search classes (compiled from /root/reference/src where they lie).
//
// The reference `abawaca` binary prints its "Best separation" line with 6
// significant digits and never instantiates ClusterSeparatorSplitScafs
// (/root/reference/src/abawaca.cpp:109-110).  This harness runs the reference's
// own ScafDpData / ClusterData / SCGdb / ClusterSeparator* objects through the
// same breadth-first work list (clusters in ascending id, children numbered at
// split time, cluster1 first -- abawaca.cpp:98-197) and dumps every evaluated
// cluster with full-precision doubles, so the flat-array oracle and the CUDA
// path can be pinned bit-for-bit, for both strategies.
//
// usage: ref_search <names> <fasta> <info> <lrn> <gene2scg|-> <scg.list> <sensspec|splitscafs> <nthreads> <out.tsv>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "ScafDpData.h"
#include "ClusterData.h"
#include "SCGdb.h"
#include "Cluster.h"
#include "ClusterQuality.h"
#include "Semaphore.h"
#include "ClusterSeparatorBySensitivitySpecificity.h"
#include "ClusterSeparatorSplitScafs.h"
#ifdef ABW_GPU_ADAPTER
// same driver, but every separate() call goes to libabawaca_b200.so through the drop-in adapter (ref_gpu_adapter.h)
#include "ref_gpu_adapter.h"
#endif

int main(int argc, char** argv)
{
	if(argc != 10) {
		fprintf(stderr, "usage: %s <names> <fasta> <info> <lrn> <gene2scg|-> <scg.list> <sensspec|splitscafs> <nthreads> <out.tsv>\n", argv[0]);
		return 2;
	}
	std::string gene2scg = (strcmp(argv[5], "-") == 0)? "" : argv[5];
	bool splitscafs = (strcmp(argv[7], "splitscafs") == 0);
	int nthreads = atoi(argv[8]);
	FILE* out = fopen(argv[9], "w");
	if(out == NULL) {
		perror(argv[9]);
		return 2;
	}

	ScafDpData	scaf_db(argv[1], argv[2], argv[3]);
	ClusterData	all_data(argv[4], scaf_db);
	SCGdb		scg_db(scaf_db, gene2scg, argv[6]);
	Semaphore	sem(nthreads);
	ClusterQuality	cq(scaf_db, scg_db);

	fprintf(out, "#N\t%lu\tS\t%lu\tD\t%lu\n", scaf_db.ndps(), scaf_db.nscafs(), all_data.ndimensions());
#ifdef ABW_GPU_ADAPTER
	abw_ctx* gpu = NULL;
	if(abw_ctx_create(0, &gpu) != ABW_OK) {
		fprintf(stderr, "no usable CUDA device\n");
		return 3;
	}
#endif

	std::map<size_t, Cluster*> work;
	std::map<size_t, size_t> scaf_bin;
	size_t next_id = 1;

	Cluster* root = new Cluster(scaf_db);
	for(size_t dp = 1; dp <= scaf_db.ndps(); dp++)
		*root += dp;
	for(size_t s = 1; s <= scaf_db.nscafs(); s++)
		root->add_assigned_scaf(s);
	work[next_id++] = root;

	// wall clock of the work list alone (objects are built from the text files above): printed on stderr as SEARCH_SECONDS for bench.py's compute-only baseline
	const std::chrono::steady_clock::time_point t_search0 = std::chrono::steady_clock::now();
	while(!work.empty()) {
		size_t id = work.begin()->first;
		Cluster* cur = work.begin()->second;
		work.erase(work.begin());
		ClusterData* sub = new ClusterData(all_data, cur->get_dps());

		ClusterSeparator* sep;
#ifdef ABW_GPU_ADAPTER
		if(splitscafs)
			sep = new ClusterSeparatorGPUSplitScafs(gpu, scaf_db, scg_db, *sub, sem);
		else
			sep = new ClusterSeparatorGPUSensSpec(gpu, scaf_db, scg_db, *sub, sem);
#else
		if(splitscafs)
			sep = new ClusterSeparatorSplitScafs(scaf_db, scg_db, *sub, sem);
		else
			sep = new ClusterSeparatorBySensitivitySpecificity(scaf_db, scg_db, *sub, sem);
#endif
		bool split = sep->separate();

		double a = 0, b = 0;
		if(splitscafs) {
			a = dynamic_cast<ClusterSeparatorSplitScafs*>(sep)->get_best_split_scaf_ratio();
		}
		else {
			a = dynamic_cast<ClusterSeparatorBySensitivitySpecificity*>(sep)->get_best_sensitivity();
			b = dynamic_cast<ClusterSeparatorBySensitivitySpecificity*>(sep)->get_best_specificity();
		}
		// dimension is size_t(-1) and value -1 when nothing was found / the split was rejected
		fprintf(out, "C\t%lu\t%lu\t%lu\t%d\t%ld\t%.17g\t%.17g\t%.17g", id, cur->ndps(), cur->nassigned_scafs(), split? 1 : 0,
			(long)sep->get_separating_dimension(), sep->get_separating_value(), a, b);
		if(split) {
			Cluster* c1 = sep->get_cluster1();
			Cluster* c2 = sep->get_cluster2();
			fprintf(out, "\t%lu\t%lu\t%lu\t%lu\t%lu\t%lu\t%lu\t%lu\n", next_id, c1->ndps(), c1->nassigned_scafs(), sep->get_raw_dps_cluster1().size(),
				next_id+1, c2->ndps(), c2->nassigned_scafs(), sep->get_raw_dps_cluster2().size());
			work[next_id++] = c1;
			work[next_id++] = c2;
		}
		else {
			double nunique = 0, avg = 0;
			cq.scg(*cur, nunique, avg);
			fprintf(out, "\tT\t%lu\t%.17g\t%.17g\n", cq.total_size(*cur), nunique, avg);
			for(auto it = cur->assigned_scafs_begin(); it != cur->assigned_scafs_end(); it++)
				scaf_bin[*it] = id;
		}
		delete sub;
		delete cur;
		// the separator leaks its ClusterQuality/ClusteringResult by design (ClusterSeparator.h:38)
	}

	fprintf(stderr, "SEARCH_SECONDS %.6f\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_search0).count());
	for(size_t s = 1; s <= scaf_db.nscafs(); s++) {
		auto it = scaf_bin.find(s);
		fprintf(out, "S\t%s\t%lu\n", scaf_db.scaf_id2name(s).c_str(), (it == scaf_bin.end())? 0UL : it->second);
	}
	fclose(out);
	return 0;
}
