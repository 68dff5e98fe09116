// TEST INFRASTRUCTURE ONLY -- the drop-in adapter of INTEGRATION.md section 1, compiled against the UNMODIFIED reference headers.
//
// ClusterSeparatorGPU<Base> derives from one of the reference's own strategy classes and overrides the virtual
// ClusterSeparator::separate() (ClusterSeparator.h:44): instead of one std::thread per dimension it hands the cluster
// to libabawaca_b200.so (abw_search_create + one level of abw_search_run) and fills the reference's own result members
// (cluster1, cluster2, raw_dps_cluster1/2, best_separation).  Everything else in the reference stays untouched, so the
// reference's own work list can run on top of it; tests/test_gpu_adapter.py checks that it then produces the same dump
// as the pure reference.
#pragma once
#include <algorithm>
#include <map>
#include <stdexcept>
#include <vector>
#include "ClusterSeparatorBySensitivitySpecificity.h"
#include "ClusterSeparatorSplitScafs.h"
#include "abawaca_b200.h"

template <class Base, int STRATEGY>
class ClusterSeparatorGPU : public Base {
public:
	ClusterSeparatorGPU(abw_ctx* _ctx, const ScafDpData& scaf_db, const SCGdb& scg_db, const ClusterData& clustering_db, Semaphore& sem) :
		Base(scaf_db, scg_db, clustering_db, sem), ctx(_ctx)	{}
	bool separate() override
	{
		// same reset as ClusterSeparator::separate(), ClusterSeparator.cpp:60-67
		if(this->cluster1 != NULL) delete(this->cluster1);
		if(this->cluster2 != NULL) delete(this->cluster2);
		this->cluster1 = this->cluster2 = NULL;
		this->best_separation->reset();
		this->raw_dps_cluster1.clear();
		this->raw_dps_cluster2.clear();

		const ClusterData& db = this->clustering_db;
		const ScafDpData& sdb = this->scaf_db;
		std::vector<size_t> dps(db.datapoints().begin(), db.datapoints().end());
		std::sort(dps.begin(), dps.end());           // dp ids are handed out scaffold by scaffold (ScafDpData.cpp:91-99): sorted = grouped
		const size_t N = dps.size(), D = db.ndimensions();
		if(N == 0)
			return false;
		std::vector<uint32_t> dp2scaf(N), T;
		std::vector<uint64_t> len;
		std::vector<size_t> scaf_ids;
		for(size_t i = 0; i < N; i++) {
			size_t s = sdb.dp2scaf(dps[i]);
			if(scaf_ids.empty() || scaf_ids.back() != s) {
				scaf_ids.push_back(s);
				T.push_back((uint32_t)sdb.ndps(s));
				len.push_back(sdb.get_scaf(s)->get_seq_size());
			}
			dp2scaf[i] = (uint32_t)(scaf_ids.size() - 1);
		}
		const size_t S = scaf_ids.size();
		// SCG sets through the public tally (SCGdb.h:19): one bit per distinct name
		std::map<std::string, size_t> name2bit;
		std::vector<std::vector<size_t> > bits(S);
		for(size_t i = 0; i < S; i++) {
			std::set<size_t> one;
			one.insert(scaf_ids[i]);
			std::map<std::string, size_t> tally;
			this->scg_db.get_scg2count(one, tally);
			for(auto it = tally.begin(); it != tally.end(); it++) {
				auto nb = name2bit.insert(std::make_pair(it->first, name2bit.size())).first;
				bits[i].push_back(nb->second);
			}
		}
		const uint32_t W = (uint32_t)std::max<size_t>(1, (name2bit.size() + 63) / 64);
		std::vector<uint64_t> mask(S * W, 0);
		for(size_t i = 0; i < S; i++)
			for(size_t b : bits[i])
				mask[i * W + b / 64] |= 1ull << (b % 64);
		std::vector<double> values(D * N);
		for(size_t d = 1; d <= D; d++) {
			const Dimension& dim = db.get_dimension(d);
			for(size_t i = 0; i < N; i++)
				values[(d - 1) * N + i] = dim.get_value(dps[i]);
		}
		abw_params prm;
		abw_default_params(&prm);
		prm.min_reported_score = 0;                  // keep best_separation of clusters that do not split, as the reference does
		abw_search* search = NULL;
		if(abw_search_create(ctx, values.data(), 0, ABW_LAYOUT_COLMAJOR, N, N, NULL, N, (uint32_t)D, dp2scaf.data(), (uint32_t)S, T.data(), len.data(), mask.data(), W,
		                     &prm, STRATEGY, &search) != ABW_OK)
			throw std::invalid_argument(abw_last_error(ctx));
		abw_search_set_max_levels(search, 1);
		abw_cluster_rec rec[4];
		uint32_t nrec = 0;
		std::vector<uint32_t> dp2cluster(N), scaf2cluster(S);
		int rc = abw_search_run(ctx, search, rec, 4, &nrec, dp2cluster.data(), scaf2cluster.data());
		abw_search_destroy(search);
		if(rc != ABW_OK || nrec < 1)
			throw std::runtime_error(abw_last_error(ctx));
		if(rec[0].best.found) {
			typename Base::ClusteringResult found(rec[0].best.dim, rec[0].best.value, rec[0].best.a, rec[0].best.b);
			this->best_separation->copy(found);
		}
		if(!rec[0].split)
			return false;
		this->cluster1 = new Cluster(sdb);
		this->cluster2 = new Cluster(sdb);
		for(size_t i = 0; i < N; i++)
			*(dp2cluster[i] == rec[0].child1? this->cluster1 : this->cluster2) += dps[i];
		for(size_t i = 0; i < S; i++) {
			if(scaf2cluster[i] == rec[0].child1) this->cluster1->add_assigned_scaf(scaf_ids[i]);
			else if(scaf2cluster[i] == rec[0].child2) this->cluster2->add_assigned_scaf(scaf_ids[i]);
		}
		// raw sides before re-homing (reconstruct_best_clusters, ClusterSeparator.cpp:36-53)
		const Dimension& win = db.get_dimension(rec[0].best.dim);
		std::vector<size_t> low, high;
		for(size_t i = 0; i < N; i++)
			(win.get_value(dps[i]) <= rec[0].best.value? low : high).push_back(dps[i]);
		const bool swapped = high.size() < low.size();
		this->raw_dps_cluster1.insert((swapped? high : low).begin(), (swapped? high : low).end());
		this->raw_dps_cluster2.insert((swapped? low : high).begin(), (swapped? low : high).end());
		return true;
	}
protected:
	void separate_dimension(size_t) override	{}   // never called: the per-dimension threads are what the device replaces
	abw_ctx* ctx;
};

typedef ClusterSeparatorGPU<ClusterSeparatorBySensitivitySpecificity, ABW_SENS_SPEC> ClusterSeparatorGPUSensSpec;
typedef ClusterSeparatorGPU<ClusterSeparatorSplitScafs, ABW_SPLIT_SCAFS> ClusterSeparatorGPUSplitScafs;
