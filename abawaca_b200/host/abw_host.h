// Host-side helpers shared by the two command-line programs (abawaca-build, abawaca) of abawaca_b200.
// Everything here is parsing and formatting at the on-disk boundary; all computation
// goes through the C ABI of libabawaca_b200.so.  Written from the format descriptions, not from the reference code.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cctype>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include <glob.h>
#include <sys/stat.h>
#include "../../include/abawaca_b200.h"

namespace abwh {

struct FastaRecord { std::string id, desc, seq; };

// FASTA as SeqIORead_fasta<DNASequence>::next_seq reads it (SeqIORead_fasta.h:51-103): blank lines skipped, id = first token after
// '>', the rest of the header is the description, sequence lines trimmed on both sides and concatenated.
inline std::vector<FastaRecord> read_fasta(const std::string& path)
{
	FILE* fp = fopen(path.c_str(), "r");
	if(!fp)
		throw std::runtime_error("Failed to open file " + path);
	std::vector<FastaRecord> out;
	char* line = nullptr;
	size_t cap = 0;
	ssize_t n;
	while((n = getline(&line, &cap, fp)) >= 0) {
		char *b = line, *e = line + n;
		while(b < e && isspace((unsigned char)*b)) b++;
		while(e > b && isspace((unsigned char)e[-1])) e--;
		if(b == e)
			continue;
		if(*b == '>') {
			FastaRecord r;
			char* p = b + 1;
			char* q = p;
			while(q < e && !isspace((unsigned char)*q)) q++;
			r.id.assign(p, q);
			while(q < e && isspace((unsigned char)*q)) q++;
			r.desc.assign(q, e);
			if(r.id.empty()) {
				free(line);
				fclose(fp);
				throw std::runtime_error("Was expecting a header line for the next sequence in fasta file " + path);
			}
			out.push_back(r);
		}
		else {
			if(out.empty()) {
				free(line);
				fclose(fp);
				throw std::runtime_error("Was expecting a header line for the next sequence in fasta file " + path);
			}
			out.back().seq.append(b, e);
		}
	}
	free(line);
	fclose(fp);
	return out;
}

inline bool directory_exists(const std::string& d)
{
	struct stat st;
	return stat(d.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

inline bool file_readable(const std::string& f)
{
	FILE* fp = fopen(f.c_str(), "r");
	if(!fp)
		return false;
	fclose(fp);
	return true;
}

inline std::vector<std::string> glob_files(const std::string& pattern)
{
	glob_t g;
	std::vector<std::string> out;
	if(glob(pattern.c_str(), GLOB_TILDE, nullptr, &g) == 0)
		for(size_t i = 0; i < g.gl_pathc; i++)
			out.push_back(g.gl_pathv[i]);
	globfree(&g);
	return out;
}

// num_snps() of a SAM line as ReadMapping::determine_snps counts them (ReadMapping.cpp:78-185): without an MD:Z field 0; otherwise one
// per mismatch letter, one per deleted reference base (after '^'), plus one per insertion operation of the CIGAR string.
inline uint32_t count_snps(const std::string& cigar, const std::string& mdz /* starts with MD:Z: or empty */)
{
	if(mdz.empty())
		return 0;
	if(mdz.compare(0, 5, "MD:Z:") != 0)
		throw std::runtime_error("Illegal SNPs description, expected MD:Z: at the beginning: " + mdz);
	uint32_t n = 0;
	size_t i = 5;
	while(i < mdz.size() && isdigit((unsigned char)mdz[i])) i++;
	while(i < mdz.size()) {
		if(mdz[i] == '^') {
			i++;
			while(i < mdz.size() && !isdigit((unsigned char)mdz[i])) { n++; i++; }
		}
		else if(mdz[i] >= 'A' && mdz[i] <= 'Z') {
			n++;
			i++;
		}
		else
			throw std::runtime_error("Illegal SNPs description: " + mdz);
		if(i >= mdz.size() || !isdigit((unsigned char)mdz[i]))
			throw std::runtime_error("Illegal SNPs description: " + mdz);
		while(i < mdz.size() && isdigit((unsigned char)mdz[i])) i++;
	}
	for(char c : cigar)
		if(c == 'I')
			n++;
	return n;
}

#define ABWH_CHECK(ctx, call)                                                                           \
	do {                                                                                                \
		int rc__ = (call);                                                                              \
		if(rc__ != ABW_OK)                                                                              \
			throw std::runtime_error(std::string(#call) + " failed: " + abw_last_error(ctx));           \
	} while(0)

}  // namespace abwh
