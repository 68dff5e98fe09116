// TEST / BASELINE INFRASTRUCTURE ONLY -- compute-only timing of the UNMODIFIED reference feature stage.
//
// The reference ships an in-memory mapping source, Bio::VectorReader (ReadMappingReader.h:76-86, .cpp:16-32).  This driver parses the FASTA and SAM
// text with the reference's own readers FIRST (untimed), then times exactly what abawaca-build computes from the parsed objects:
//   phase "scaf"  : Scaf::Scaf for every sequence (abawaca-build.cpp:198-228) with Scaf_segment::Scaf_segment inside (:103-174): windows + k-mer signature
//   phase "reads" : per sample, VectorReader::next_mapping -> the read filter (:546-550) -> Scaf::add_mapped_read (:231-244, :177-185)
// so that a text-free CPU figure stands beside the GPU path, whose timed region starts from parsed records as well.
//
// usage: ref_compute <assembly.fa> <sample.sam> [<sample.sam> ...]        prints one JSON line
#include <chrono>
#include <cstdlib>

#define main abawaca_build_reference_main
#include "abawaca-build.cpp"  // resolved through -I/root/reference/src
#undef main

int main(int argc, char** argv)
{
	if(argc < 3) {
		fprintf(stderr, "usage: %s <assembly.fa> <sample.sam> [<sample.sam> ...]\n", argv[0]);
		return 2;
	}
	typedef std::chrono::steady_clock clk;
	// ---- untimed: text -> objects, with the reference's readers
	vector<Bio::DNASequence*> seqs;
	{
		Bio::SeqIORead_fasta<Bio::DNASequence> in(argv[1]);
		Bio::DNASequence* seq = NULL;
		while((seq = in.next_seq()) != NULL)
			seqs.push_back(seq);
	}
	const int nsam = argc - 2;
	vector<vector<const Bio::ReadMapping*> > mappings(nsam);
	size_t nreads = 0;
	for(int f = 0; f < nsam; f++) {
		Bio::SAMReader reader(argv[2 + f]);
		while(reader.good()) {
			Bio::ReadMapping* m = reader.next_mapping();
			if(m == NULL)
				break;
			mappings[f].push_back(m);
		}
		nreads += mappings[f].size();
	}
	// ---- timed: windows and k-mer signatures
	const clk::time_point t0 = clk::now();
	map<string, Scaf*> scafs;
	size_t ndps = 0;
	for(size_t i = 0; i < seqs.size(); i++) {
		Scaf* s = new Scaf(*seqs[i], 2000);
		scafs[seqs[i]->display_id()] = s;
		ndps += s->ndps();
	}
	const clk::time_point t1 = clk::now();
	// ---- timed: coverage from the in-memory mappings
	size_t accepted = 0;
	for(int f = 0; f < nsam; f++) {
		string sam = argv[2 + f];
		dimension2index.insert(pair<string, size_t>(sam, dimension_order.size()));
		dimension_order.push_back(sam);
		const size_t dimension = dimension2index[sam];
		if(f == 0)
			this_sample = dimension;
		Bio::VectorReader reader(mappings[f]);
		while(reader.good()) {
			Bio::ReadMappingPtr mapping(reader.next_mapping());
			if(mapping == NULL)
				break;
			if(mapping->unmapped() || (mapping->num_snps() > 15) || mapping->multiple_hits())
				continue;
			auto mit = scafs.find(mapping->ref_name());
			if(mit != scafs.end()) {
				mit->second->add_mapped_read(*mapping, dimension);
				accepted++;
			}
		}
	}
	const clk::time_point t2 = clk::now();
	printf("{\"scaffolds\": %lu, \"windows\": %lu, \"reads\": %lu, \"accepted\": %lu, \"samples\": %d, \"scaf_s\": %.6f, \"reads_s\": %.6f}\n", (unsigned long)seqs.size(),
	       (unsigned long)ndps, (unsigned long)nreads, (unsigned long)accepted, nsam, std::chrono::duration<double>(t1 - t0).count(),
	       std::chrono::duration<double>(t2 - t1).count());
	return 0;
}
