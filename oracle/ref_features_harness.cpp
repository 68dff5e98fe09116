// TEST INFRASTRUCTURE ONLY -- function-level driver around the UNMODIFIED reference
// feature-stage classes.  `Scaf` / `Scaf_segment` are file-local to
// /root/reference/src/abawaca-build.cpp (:34-62), so that translation unit is
// compiled where it lies with its `main` renamed, and this driver calls the
// classes directly.  It dumps UN-truncated doubles (the .lrn file written by
// the real binary only holds trunc(1000*x)/1000), which pins the k-mer
// arithmetic (abawaca-build.cpp:103-174) and the order-dependent coverage sum
// (:177-185, :231-244) more tightly than the file does.
//
// usage: ref_features <assembly.fa> <out.tsv> [<sample.sam> ...]
//   the FIRST sam file plays the role of `-c` (this_sample)
#include <cstdlib>

#define main abawaca_build_reference_main
#include "abawaca-build.cpp"  // resolved through -I/root/reference/src
#undef main

int main(int argc, char** argv)
{
	if(argc < 3) {
		fprintf(stderr, "usage: %s <assembly.fa> <out.tsv> [<sample.sam> ...]\n", argv[0]);
		return 2;
	}
	FILE* out = fopen(argv[2], "w");
	if(out == NULL) {
		perror(argv[2]);
		return 2;
	}

	// scaffolds keyed by name, as abawaca-build.cpp:482-490
	map<string, Scaf*> scafs;
	Bio::SeqIORead_fasta<Bio::DNASequence> in(argv[1]);
	Bio::DNASequence* seq = NULL;
	while((seq = in.next_seq()) != NULL)
		scafs[seq->display_id()] = new Scaf(*seq, 2000);

	size_t n_kmer_dims = dimension_order.size();
	int nsam = argc - 3;
	for(int f = 0; f < nsam; f++) {
		string sam = argv[3+f];
		dimension2index.insert(pair<string, size_t>(sam, dimension_order.size()));
		dimension_order.push_back(sam);
		size_t dimension = dimension2index[sam];
		if(f == 0)
			this_sample = dimension;
		Bio::SAMReader reader(sam);
		while(reader.good()) {
			Bio::ReadMappingPtr mapping(reader.next_mapping());
			if(mapping == NULL)
				break;
			if(mapping->unmapped() || (mapping->num_snps() > 15) || mapping->multiple_hits())
				continue;
			auto mit = scafs.find(mapping->ref_name());
			if(mit != scafs.end())
				mit->second->add_mapped_read(*mapping, dimension);
		}
	}

	fprintf(out, "#dims");
	for(size_t d = 0; d < n_kmer_dims; d++)
		fprintf(out, "\t%s", dimension_order[d].c_str());
	fprintf(out, "\n");
	for(auto it = scafs.begin(); it != scafs.end(); it++) {
		Scaf* s = it->second;
		fprintf(out, "SCAF\t%s\t%lu\t%lu\t%.17g\t%.17g\t%lu\n", s->get_display_id().c_str(), s->get_seq().size(), s->ndps(), s->cvg(), s->gc(), s->Ns());
		for(auto sit = s->begin(); sit != s->end(); sit++) {
			fprintf(out, "SEG\t%s\t%s", (*sit)->get_display_id().c_str(), (*sit)->get_desc().c_str());
			// checked accessor (:39) -- returns 0 past the end instead of the unchecked read of :188-193
			for(size_t d = 0; d < n_kmer_dims + nsam; d++)
				fprintf(out, "\t%.17g", (*sit)->get_dimension(d));
			fprintf(out, "\n");
		}
	}
	fclose(out);
	return 0;
}
