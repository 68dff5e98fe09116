// abawaca-build (B200): same command line and same output files as the reference program
// (/root/reference/src/abawaca-build.cpp main, :434-634), with the feature stage on the device:
//   SeqIORead_fasta::next_seq       -> abw_fasta_scan / abw_fasta_pack (the FASTA text goes to the device as it is)
//   Scaf::Scaf                      -> abw_segment
//   Scaf_segment::Scaf_segment      -> abw_kmer_features
//   Scaf::add_mapped_read           -> abw_coverage (one call per SAM file, reads in file order)
//   SAMReader / ReadMapping(line)   -> abw_parse_sam (SAM text goes to the device in 64 MB chunks cut at line boundaries)
// The text writers stay on the host.  Not ported (out of the hot path):
// paired-end link counting (abawaca.links is written empty: `abawaca` only checks that it exists) and the calls to
// the external gene predictor / SCG script.
#include "abw_host.h"
#include <algorithm>
#include <ctime>
#include <iostream>

using namespace abwh;

static std::string stamp()
{
	static time_t t0 = time(nullptr);
	unsigned t = (unsigned)(time(nullptr) - t0);
	char b[32];
	snprintf(b, sizeof(b), "[%02u:%02u:%02u]\t", t / 3600, (t % 3600) / 60, t % 60);
	return b;
}

static void usage(const char* prog)
{
	std::cerr << std::endl << "Usage: " << prog << " -f <fasta-file> -o <out-directory> -s <sam-files-glob> -c <source-sam-file>" << std::endl << std::endl;
}

// the 180 canonical 1..4-mer names in .lrn order: first of {mer, reverse complement} in lexicographic enumeration (abawaca-build.cpp:75-100)
static std::vector<std::string> kmer_names()
{
	std::vector<std::string> names;
	std::map<std::string, size_t> seen;
	for(int k = 1; k <= 4; k++)
		for(int code = 0; code < (1 << (2 * k)); code++) {
			std::string mer(k, 'A'), rc(k, 'A');
			for(int i = 0; i < k; i++) {
				int b = (code >> (2 * (k - 1 - i))) & 3;
				mer[i] = "ACGT"[b];
				rc[k - 1 - i] = "TGCA"[b];
			}
			auto it = seen.find(rc);
			if(it == seen.end()) {
				seen[mer] = names.size();
				names.push_back(mer);
			}
			else
				seen[mer] = it->second;
		}
	return names;
}

int main(int argc, const char* argv[])
{
	const uint32_t window_size = 2000, max_snps = 15;   // abawaca-build.cpp:436-437
	if(argc < 5) {
		usage(argv[0]);
		return -1;
	}
	std::string assembly_file, out_directory, sam_glob, this_sample_str;
	for(int i = 1; i < argc; i += 2) {
		const char* v = (i + 1 < argc)? argv[i + 1] : "";
		if(!strcmp(argv[i], "-f")) assembly_file = v;
		else if(!strcmp(argv[i], "-o")) out_directory = v;
		else if(!strcmp(argv[i], "-s")) sam_glob = v;
		else if(!strcmp(argv[i], "-c")) this_sample_str = v;
		else {
			std::cerr << std::endl << "Unknown option " << argv[i] << std::endl << std::endl;
			return -1;
		}
	}
	if(assembly_file.empty()) { std::cerr << std::endl << "Error: assembly file was not specified (-f)" << std::endl << std::endl; return -1; }
	if(out_directory.empty()) { std::cerr << std::endl << "Error: output directory was not specified (-o)" << std::endl << std::endl; return -1; }
	if(sam_glob.empty()) { std::cerr << std::endl << "Error: glob path for sam files was not specified (-s)" << std::endl << std::endl; return -1; }
	if(this_sample_str.empty()) { std::cerr << std::endl << "Error: path for sam file for the assembly's reads was not specified (-f)" << std::endl << std::endl; return -1; }

	try {
		std::cerr << stamp() << "Reading assembly file (" << assembly_file << ")" << std::endl;
		abw_ctx* ctx = nullptr;
		if(abw_ctx_create(0, &ctx) != ABW_OK) {
			std::cerr << "Error: no usable CUDA device (abawaca_b200 has no CPU path)" << std::endl;
			return -1;
		}
		// the FASTA text goes to the device as it is: records are indexed there (abw_fasta_scan), the host only reads the names
		std::string fasta_text;
		{
			FILE* fp = fopen(assembly_file.c_str(), "rb");
			if(!fp)
				throw std::runtime_error("Failed to open file " + assembly_file);
			fseek(fp, 0, SEEK_END);
			const long sz = ftell(fp);
			rewind(fp);
			fasta_text.resize((size_t)std::max<long>(sz, 0));
			if(sz > 0 && fread(&fasta_text[0], 1, (size_t)sz, fp) != (size_t)sz)
				throw std::runtime_error("Failed to read file " + assembly_file);
			fclose(fp);
		}
		abw_fasta* fa = nullptr;
		if(abw_fasta_scan(ctx, fasta_text.data(), fasta_text.size(), 0, &fa) != ABW_OK)
			throw std::runtime_error(std::string(abw_last_error(ctx)) + " (" + assembly_file + ")");
		const uint64_t nrec = abw_fasta_count(fa);
		std::vector<uint64_t> id_off(nrec), rec_len(nrec);
		std::vector<uint32_t> id_len(nrec);
		ABWH_CHECK(ctx, abw_fasta_get(ctx, fa, id_off.data(), id_len.data(), rec_len.data()));
		// scaffolds in name order; the first record of a name wins (std::map::insert, abawaca-build.cpp:482-490)
		std::map<std::string, size_t> by_name;
		for(size_t i = 0; i < nrec; i++) {
			if(rec_len[i] == 0)
				std::cerr << "Warning: sequence " << fasta_text.substr(id_off[i], id_len[i]) << " is empty" << std::endl;      // SeqIORead_fasta.h:99-101
			by_name.insert(std::make_pair(fasta_text.substr(id_off[i], id_len[i]), i));
		}
		const uint32_t nscaf = (uint32_t)by_name.size();
		std::vector<std::string> scaf_name;
		std::vector<uint64_t> scaf_len;
		std::vector<uint32_t> order;
		for(auto& kv : by_name) {
			scaf_name.push_back(kv.first);
			scaf_len.push_back(rec_len[kv.second]);
			order.push_back((uint32_t)kv.second);
		}
		abw_seqset* ss = nullptr;
		int rc = abw_fasta_pack(ctx, fa, order.data(), nscaf, &ss);
		abw_fasta_destroy(fa);
		std::string().swap(fasta_text);
		if(rc != ABW_OK) {
			std::cerr << "Fatal error, attempted to initialize DNAString with illegal string: " << abw_last_error(ctx) << std::endl;
			return -1;
		}
		abw_segments* sg = nullptr;
		ABWH_CHECK(ctx, abw_segment(ctx, ss, window_size, &sg));
		const uint64_t ndps = abw_segments_count(sg);

		if(!directory_exists(out_directory)) {
			std::string cmd = "mkdir " + out_directory;
			if(system(cmd.c_str()) != 0)
				std::cerr << "Warning: could not create " << out_directory << std::endl;
		}
		const std::string names_file = out_directory + "/abawaca.names", lrn_file = out_directory + "/abawaca.lrn", info_file = out_directory + "/abawaca.info",
		                  links_file = out_directory + "/abawaca.links";
		std::cerr << stamp() << "Creating .lrn (" << lrn_file << ") and .names (" << names_file << ") files" << std::endl;

		const std::vector<std::string> sam_files = glob_files(sam_glob);
		const uint64_t ld = 179 + sam_files.size();
		double* d_rows = nullptr;
		uint64_t* d_nbps = nullptr;
		ABWH_CHECK(ctx, abw_device_alloc(ctx, std::max<uint64_t>(ndps * ld, 1) * sizeof(double), (void**)&d_rows));
		ABWH_CHECK(ctx, abw_device_alloc(ctx, std::max<uint32_t>(nscaf, 1) * sizeof(uint64_t), (void**)&d_nbps));
		ABWH_CHECK(ctx, abw_memset_device(ctx, d_nbps, 0, std::max<uint32_t>(nscaf, 1) * sizeof(uint64_t)));
		ABWH_CHECK(ctx, abw_kmer_features(ctx, ss, sg, ABW_FEAT_TRUNC3, 1, d_rows, ld, 0));

		// the scaffold names as a device-side table: SAM text is parsed on the device (abw_parse_sam)
		abw_names* names = nullptr;
		{
			std::string blob;
			std::vector<uint64_t> noff(1, 0);
			for(auto& kv : by_name) {
				blob += kv.first;
				noff.push_back(blob.size());
			}
			ABWH_CHECK(ctx, abw_names_create(ctx, blob.data(), noff.data(), nscaf, &names));
		}
		const size_t CHUNK = (size_t)64 << 20;
		std::vector<char> buf(CHUNK + 1);
		for(size_t j = 0; j < sam_files.size(); j++) {
			std::cerr << stamp() << "Reading SAM file " << sam_files[j] << std::endl;
			const bool is_this_sample = (sam_files[j] == this_sample_str);
			FILE* fp = fopen(sam_files[j].c_str(), "rb");
			if(!fp)
				throw std::runtime_error("Failed to open SAM file " + sam_files[j]);
			// a record line has at least 11 fields, i.e. at least 10 tabs and a newline: the file size bounds the number of records
			fseek(fp, 0, SEEK_END);
			const uint64_t nlines = (uint64_t)ftell(fp) / 11 + 2;
			rewind(fp);
			abw_read* d_reads = nullptr;
			ABWH_CHECK(ctx, abw_device_alloc(ctx, std::max<uint64_t>(nlines, 1) * sizeof(abw_read), (void**)&d_reads));
			uint64_t nreads = 0;
			size_t have = 0;                                       // bytes of an unfinished line carried over from the previous chunk
			for(;;) {
				if(have == buf.size() - 1)
					buf.resize(2 * buf.size());                    // a line longer than the chunk
				const size_t n = fread(buf.data() + have, 1, buf.size() - 1 - have, fp);
				const size_t total = have + n;
				if(total == 0)
					break;
				size_t cut = total;
				if(n > 0) {                                        // not at the end of the file: cut after the last complete line
					while(cut > 0 && buf[cut - 1] != '\n')
						cut--;
					if(cut == 0) {
						have = total;
						continue;
					}
				}
				uint64_t got = 0;
				int prc = abw_parse_sam(ctx, names, buf.data(), cut, 0, d_reads + nreads, nlines - nreads, &got);
				if(prc == ABW_ERR_ILLEGAL_DNA) {
					std::cerr << "Fatal error, attempted to initialize DNAString with illegal string: " << abw_last_error(ctx) << std::endl;
					return -1;
				}
				if(prc != ABW_OK)
					throw std::runtime_error(std::string("abw_parse_sam failed on ") + sam_files[j] + ": " + abw_last_error(ctx));
				nreads += got;
				have = total - cut;
				memmove(buf.data(), buf.data() + cut, have);
				if(n == 0)
					break;
			}
			fclose(fp);
			ABWH_CHECK(ctx, abw_coverage(ctx, sg, d_reads, nreads, 1, max_snps, ABW_FEAT_TRUNC3, d_rows, ld, (uint32_t)(179 + j), is_this_sample? d_nbps : nullptr));
			ABWH_CHECK(ctx, abw_device_free(ctx, d_reads));
			if(is_this_sample) {
				FILE* fl = fopen(links_file.c_str(), "w");     // scaffold-end links are not computed; the file must exist
				if(fl) fclose(fl);
			}
		}
		abw_names_destroy(names);

		// fetch everything the three text files need
		std::vector<double> rows((size_t)ndps * ld);
		std::vector<uint32_t> seg_scaf(ndps);
		std::vector<uint64_t> seg_start(ndps), seg_end(ndps), seg_nonN(ndps), nbps(nscaf), nN(nscaf), nGC(nscaf);
		if(ndps)
			ABWH_CHECK(ctx, abw_copy_to_host(ctx, rows.data(), d_rows, rows.size() * sizeof(double)));
		ABWH_CHECK(ctx, abw_segments_get(ctx, sg, nullptr, seg_scaf.data(), seg_start.data(), seg_end.data(), seg_nonN.data()));
		if(nscaf)
			ABWH_CHECK(ctx, abw_copy_to_host(ctx, nbps.data(), d_nbps, nscaf * sizeof(uint64_t)));
		ABWH_CHECK(ctx, abw_seqset_stats(ctx, ss, nN.data(), nGC.data()));

		FILE* fout = fopen(names_file.c_str(), "w");
		FILE* flrn = fopen(lrn_file.c_str(), "w");
		FILE* finfo = fopen(info_file.c_str(), "w");
		if(!fout) { std::cerr << "Could not write to " << names_file << std::endl << std::endl; return -1; }
		if(!flrn) { std::cerr << "Could not write to " << lrn_file << std::endl << std::endl; return -1; }
		if(!finfo) { std::cerr << "Could not write to " << info_file << std::endl << std::endl; return -1; }
		fprintf(fout, "%c %lu\n", '%', (unsigned long)ndps);
		fprintf(flrn, "%c %lu\n", '%', (unsigned long)ndps);
		fprintf(flrn, "%c %lu\n", '%', (unsigned long)(180 + sam_files.size()));     // 180 + samples: the unwritten "A" and the key column cancel
		fprintf(flrn, "%c 9", '%');
		for(uint64_t c = 0; c < ld; c++) fprintf(flrn, "\t1");
		fprintf(flrn, "\n%c Key", '%');
		const std::vector<std::string> kn = kmer_names();
		for(size_t d = 1; d < kn.size(); d++) fprintf(flrn, "\t%s", kn[d].c_str());
		for(auto& sf : sam_files) fprintf(flrn, "\t%s", sf.c_str());
		fprintf(flrn, "\n");
		uint64_t g = 0;
		for(uint32_t s = 0; s < nscaf; s++) {
			const std::string& id = scaf_name[s];
			const double len = (double)scaf_len[s];
			const double cvg = (double)nbps[s] / len;
			const double denom = len - (double)nN[s];
			const double gc = (denom == 0)? 0 : (double)nGC[s] / denom;
			fprintf(finfo, "%s\t%lu\t%.3lf\t%.3lf\t%lu\n", id.c_str(), (unsigned long)scaf_len[s], int(1000.0 * cvg) / 1000.0, int(1000.0 * gc) / 1000.0,
			        (unsigned long)nN[s]);
			unsigned long k = 0;
			for(; g < ndps && seg_scaf[g] == s; g++) {
				k++;
				fprintf(fout, "%lu\t%s_%lu\t%s:(%lu, %lu), %lu/%lu non-Ns bps\n", (unsigned long)(g + 1), id.c_str(), k, id.c_str(), (unsigned long)seg_start[g],
				        (unsigned long)seg_end[g], (unsigned long)seg_nonN[g], (unsigned long)(seg_end[g] - seg_start[g] + 1));
				fprintf(flrn, "%lu", (unsigned long)(g + 1));
				for(uint64_t c = 0; c < ld; c++)
					fprintf(flrn, "\t%.3lf", rows[g * ld + c]);
				fprintf(flrn, "\n");
			}
		}
		fclose(fout);
		fclose(flrn);
		fclose(finfo);

		std::cerr << stamp() << "Predicting genes and SCGs (skipped: external tools are not part of this build; pass -c <gene2scg> to abawaca)" << std::endl;
		size_t l = assembly_file.rfind('/');
		const std::string out_name = (l == std::string::npos)? assembly_file : assembly_file.substr(l + 1);
		const std::string scg_file = out_directory + "/" + out_name + ".proteins.faa.bacteria.scg";
		const std::string data = out_directory + "/data.txt";
		FILE* fp = fopen(data.c_str(), "w");
		if(!fp) { std::cerr << "Could not write to " << data << std::endl << std::endl; return -1; }
		{
		// the reference's prepare_scgs leaves this file behind (abawaca-build.cpp:401-431); without the external gene caller it is created empty so that
		// the two-step command line (abawaca-build, then abawaca -u DIR -o OUT) runs: no SCG information unless abawaca gets -c <gene2scg>
		FILE* fs = fopen(scg_file.c_str(), "a");
		if(fs) fclose(fs);
	}
	fprintf(fp, "SCG\t%s\n", scg_file.c_str());
		fprintf(fp, "Links\t%s\n", links_file.c_str());
		fprintf(fp, "Info\t%s\n", info_file.c_str());
		fprintf(fp, "Names\t%s\n", names_file.c_str());
		fprintf(fp, "Lrn\t%s\n", lrn_file.c_str());
		fprintf(fp, "Assembly\t%s\n", assembly_file.c_str());
		fclose(fp);

		abw_device_free(ctx, d_rows);
		abw_device_free(ctx, d_nbps);
		abw_segments_destroy(sg);
		abw_seqset_destroy(ss);
		abw_ctx_destroy(ctx);
		std::cerr << stamp() << "Finished successfully" << std::endl;
	}
	catch(std::exception& e) {
		std::cerr << "Fatal error: " << e.what() << std::endl;
		return -1;
	}
	return 0;
}
