// reflexiv_main.cpp -- `reflexiv` command-line driver on top of libreflexiv_cuda (C ABI only).
//
// Mirrors, for the `run` and `counter` commands:
//   bin/reflexiv:209-267                     command word, `--x [v]` (Spark) vs `-x [v]` (Reflexiv) option split
//   main/Main.java:57-79, MainOfCounter.java:58-80
//   util/Parameter.java:311-611, util/ParameterOfCounter.java:205-390   option names, defaults, exit code 0 on errors
//   pipeline/ReflexivDataFrameCounter.java:222-233   <out>/Count_<k>/part-*.csv[.gz] + _SUCCESS
//   pipeline/ReflexivDSMain.java:331-354, 706-710    <out>/part-00000 (or <out>/Assemble_<k>/ with -kmerc)
// The reference's driver is Java; a JDK is not available in this image, so the host side above the C ABI is C++
// (INTEGRATION.md shows the JNI / Panama binding a Java driver would use instead).
#include <dirent.h>
#include <glob.h>
#include <sys/stat.h>
#include <time.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/reflexiv_cuda.h"

static void info(const char* msg) {  // util/InfoDumper.java: "Reflexiv HH:mm:ss msg"
    time_t t = time(nullptr);
    struct tm tmv;
    localtime_r(&t, &tmv);
    printf("Reflexiv %02d:%02d:%02d %s\n", tmv.tm_hour, tmv.tm_min, tmv.tm_sec, msg);
    fflush(stdout);
}

static const char* HELP =
    "usage: reflexiv <run|counter> [--spark-options ignored] -fastq <glob> -outfile <dir> [-kmer 31] [-cover 2]\n"
    "       [-maxcov 10000000] [-error 8] [-clipf N] [-clipe N] [-mincontig 500] [-miniter 15] [-maxiter 150]\n"
    "       [-partition N] [-partitionredu 200] [-kmerc <Count_k csv glob>] [-infmt fmt] [-bubble] [-gzip] [-cache]\n";

struct Opt { bool has_arg; };
static std::map<std::string, Opt> run_options() {
    std::map<std::string, Opt> m;
    for (const char* n : {"fastq", "paired", "single", "inter", "fasta", "infmt", "reads", "contig", "kmerc", "outfile", "kmer", "klist",
                          "overlap", "miniter", "maxiter", "clipf", "clipe", "cover", "maxcov", "error", "minlength", "mincontig",
                          "partition", "partitionredu", "sbin", "mode"})
        m[n] = Opt{true};
    for (const char* n : {"gzip", "bubble", "stitch", "accurate", "cache", "version", "h", "help"}) m[n] = Opt{false};
    return m;
}
static std::map<std::string, Opt> counter_options() {
    std::map<std::string, Opt> m;
    for (const char* n : {"fastq", "fasta", "infmt", "reads", "outfile", "kmer", "overlap", "clipf", "clipe", "cover", "maxcov", "minlength",
                          "partition", "partitionredu"})
        m[n] = Opt{true};
    for (const char* n : {"gzip", "cache", "version", "h", "help"}) m[n] = Opt{false};
    return m;
}

static int bad_params(const std::string& why) {  // Parameter.java:601-611: message, stack trace, exit code 0
    info("Parameter settings incorrect.");
    fprintf(stderr, "%s\n", why.c_str());
    return 0;
}

static bool read_file(const std::string& path, std::string& out) {
    const bool gz = path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
    if (path.size() > 4 && path.compare(path.size() - 4, 4, ".4mc") == 0) {
        fprintf(stderr, "%s: 4mc input needs hadoop-4mc; decompress first or use gzip / plain text\n", path.c_str());
        return false;
    }
    gzFile f = gzopen(path.c_str(), "rb");  // zlib reads plain files transparently
    if (!f) return false;
    (void)gz;
    char buf[1 << 16];
    int n;
    while ((n = gzread(f, buf, sizeof(buf))) > 0) out.append(buf, (size_t)n);
    gzclose(f);
    if (!out.empty() && out.back() != '\n') out.push_back('\n');
    return true;
}

static bool read_glob(const std::string& pattern, std::string& out) {
    glob_t g;
    std::vector<std::string> paths;
    if (glob(pattern.c_str(), 0, nullptr, &g) == 0)
        for (size_t i = 0; i < g.gl_pathc; i++) paths.push_back(g.gl_pathv[i]);
    globfree(&g);
    if (paths.empty()) { fprintf(stderr, "Input path does not exist: %s\n", pattern.c_str()); return false; }
    std::sort(paths.begin(), paths.end());
    for (const std::string& p : paths) {
        struct stat st;
        if (stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) {
            std::vector<std::string> inner;
            if (DIR* d = opendir(p.c_str())) {
                while (dirent* e = readdir(d))
                    if (e->d_name[0] != '_' && e->d_name[0] != '.') inner.push_back(p + "/" + e->d_name);
                closedir(d);
            }
            std::sort(inner.begin(), inner.end());
            for (const std::string& q : inner)
                if (!read_file(q, out)) return false;
        } else if (!read_file(p, out)) return false;
    }
    return true;
}

static bool write_out(const std::string& path, const char* data, size_t n, bool gz) {
    if (gz) {
        gzFile f = gzopen((path + ".gz").c_str(), "wb");
        if (!f) return false;
        size_t off = 0;
        while (off < n) { int w = gzwrite(f, data + off, (unsigned)std::min<size_t>(n - off, 1u << 30)); if (w <= 0) { gzclose(f); return false; } off += (size_t)w; }
        return gzclose(f) == Z_OK;
    }
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite(data, 1, n, f) == n;
    return fclose(f) == 0 && ok;
}

static int fail(rfx_ctx* c, const char* what) {
    fprintf(stderr, "reflexiv: %s: %s\n", what, rfx_last_error(c));
    if (c) rfx_destroy(c);
    return 1;
}

// KmerBinarizer input: `KMER,count` or legacy `(KMER,count)`, ReflexivDSMain.java:3872-3948
static bool parse_counts(const std::string& text, int k, int minc, int maxc, std::vector<uint64_t>& keys, std::vector<uint32_t>& counts) {
    const int words = k <= 31 ? 1 : k / 32 + 1;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t end = text.find('\n', pos);
        if (end == std::string::npos) end = text.size();
        std::string line = text.substr(pos, end - pos);
        pos = end + 1;
        if (line.empty()) continue;
        if (line[0] == '(') line = line.substr(1);
        if (!line.empty() && line.back() == ')') line.pop_back();
        size_t comma = line.find(',');
        if (comma == std::string::npos || (int)comma < k) return false;
        const std::string num = line.substr(comma + 1);
        long cover = num.size() >= 10 ? 1000000000L : atol(num.c_str());
        if (cover < minc || cover > maxc) continue;  // ReflexivDSMain.java:405-412
        unsigned __int128 v = 0;
        for (int i = 0; i < k; i++) { char ch = line[i]; v = (v << 2) | (ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3); }
        if (words == 1) keys.push_back((uint64_t)v);
        else { const int res = k % 32; keys.push_back((uint64_t)(v >> (2 * res))); keys.push_back((uint64_t)v & ((res ? ((uint64_t)1 << (2 * res)) : 1) - 1)); }
        counts.push_back((uint32_t)cover);
    }
    return true;
}

int main(int argc, char** argv) {
    if (argc < 2) { fputs(HELP, stdout); return 1; }
    const std::string cmd = argv[1];
    if (cmd != "run" && cmd != "counter") {
        fprintf(stderr, "reflexiv: command '%s' is outside the GPU path (supported: run, counter)\n", cmd.c_str());
        fputs(HELP, stdout);
        return 1;
    }
    const bool counter = cmd == "counter";
    // bin/reflexiv:209-238: `--x [value]` belongs to spark-submit, `-x [value]` to Reflexiv
    std::vector<std::string> own;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        const bool next_is_value = i + 1 < argc && argv[i + 1][0] != '-';
        if (a.rfind("--", 0) == 0) { if (next_is_value) i++; continue; }
        if (a[0] == '-') { own.push_back(a); if (next_is_value) own.push_back(argv[++i]); }
    }
    info(counter ? "Reflexiv counter initiating ... " : "Reflexiv main initiating ... ");
    info("interpreting parameters.");
    const auto table = counter ? counter_options() : run_options();
    std::map<std::string, std::string> v;
    for (size_t i = 0; i < own.size(); i++) {
        std::string name = own[i].substr(own[i].find_first_not_of('-'));
        auto it = table.find(name);
        if (it == table.end()) return bad_params("Unrecognized option: " + own[i]);
        if (it->second.has_arg) {
            if (i + 1 >= own.size()) return bad_params("Missing argument for option: " + name);
            v[name] = own[++i];
        } else v[name] = "1";
    }
    if (v.count("help") || v.count("h")) { fputs(HELP, stdout); return 0; }
    if (v.count("version")) return 0;

    rfx_params p;
    rfx_params_default(&p);
    std::string err;
    auto geti = [&](const char* n, int32_t& dst, long lo, long hi) {
        if (!v.count(n)) return true;
        char* e = nullptr;
        long x = strtol(v[n].c_str(), &e, 0);
        if (!e || *e) { err = std::string("For input string: \"") + v[n] + "\""; return false; }
        if (x < lo || x > hi) { err = std::string("Parameter ") + n + " out of range"; return false; }
        dst = (int32_t)x;
        return true;
    };
    if (!geti("kmer", p.kmer_size, -2147483647L, 2147483647L) || !geti("partition", p.partitions, 0, 2147483647L) ||
        !geti("partitionredu", p.shuffle_partitions, 0, 2147483647L) || !geti("miniter", p.min_iter, 0, 2147483647L) ||
        !geti("maxiter", p.max_iter, -2147483647L, 100000) || !geti("clipf", p.front_clip, 1, 2147483647L) ||
        !geti("clipe", p.end_clip, 1, 2147483647L) || !geti("cover", p.min_kmer_coverage, 0, 2147483647L) ||
        !geti("maxcov", p.max_kmer_coverage, 0, 2147483647L) || !geti("error", p.min_error_coverage, 0, 2147483647L) ||
        !geti("mincontig", p.min_contig, 0, 2147483647L))
        return bad_params(err);
    if (v.count("bubble")) p.bubble = 0;
    const bool gz = v.count("gzip") > 0;
    const std::string infmt = v.count("infmt") ? v["infmt"] : "4mc";
    const bool from_kmer = !counter && v.count("kmerc") && !v.count("fastq");
    if (!v.count("fastq") && !from_kmer) { fputs(HELP, stdout); return 0; }  // Parameter.java:565-568
    if (!v.count("outfile")) { info("Output file not set of -outfile options"); return 0; }
    const std::string outdir = v["outfile"];
    p.counter_mode = counter ? 1 : 0;
    p.fastq_mode = counter ? (infmt == "line" ? RFX_FASTQ_LINE : RFX_FASTQ_COUNTER) : RFX_FASTQ_RUN;
    if (const char* d = getenv("REFLEXIV_DEVICE")) p.device = atoi(d);

    const std::string target = counter ? outdir + "/Count_" + std::to_string(p.kmer_size)
                                       : (from_kmer ? outdir + "/Assemble_" + std::to_string(p.kmer_size) : outdir);
    struct stat stt;
    if (!counter && stat(target.c_str(), &stt) == 0) {  // Hadoop FileAlreadyExistsException in saveAsTextFile
        fprintf(stderr, "reflexiv: output directory %s already exists\n", target.c_str());
        return 1;
    }

    info("Initiating CUDA context ...");
    rfx_ctx* c = nullptr;
    if (rfx_create(&c, &p) != RFX_OK) return fail(nullptr, "rfx_create");
    std::string text;
    if (from_kmer) {
        if (!read_glob(v["kmerc"], text)) { rfx_destroy(c); return 1; }
        std::vector<uint64_t> keys;
        std::vector<uint32_t> counts;
        if (!parse_counts(text, p.kmer_size, p.min_kmer_coverage, p.max_kmer_coverage, keys, counts)) { fprintf(stderr, "reflexiv: malformed k-mer count row\n"); rfx_destroy(c); return 1; }
        if (rfx_load_counts(c, keys.data(), counts.data(), counts.size()) != RFX_OK) return fail(c, "rfx_load_counts");
    } else {
        if (!read_glob(v["fastq"], text)) { rfx_destroy(c); return 1; }
        if (rfx_push_fastq(c, reinterpret_cast<const uint8_t*>(text.data()), text.size()) != RFX_OK) return fail(c, "rfx_push_fastq");
        std::string().swap(text);
        if (rfx_count(c) != RFX_OK) return fail(c, "rfx_count");
    }
    mkdir(outdir.c_str(), 0755);
    if (counter) {
        uint64_t nbytes = 0;
        if (rfx_counts_csv(c, nullptr, 0, &nbytes) != RFX_OK) return fail(c, "rfx_counts_csv");
        std::vector<char> csv(nbytes ? nbytes : 1);
        if (rfx_counts_csv(c, csv.data(), nbytes, &nbytes) != RFX_OK) return fail(c, "rfx_counts_csv");
        mkdir(target.c_str(), 0755);
        if (DIR* d = opendir(target.c_str())) {  // SaveMode.Overwrite
            while (dirent* e = readdir(d))
                if (e->d_name[0] != '.' || strlen(e->d_name) > 2) remove((target + "/" + e->d_name).c_str());
            closedir(d);
        }
        char name[128];
        snprintf(name, sizeof(name), "/part-00000-%08lx-%04x-%04x-%04x-%012lx-c000.csv", (unsigned long)time(nullptr) & 0xffffffffUL, rand() & 0xffff,
                 rand() & 0xffff, rand() & 0xffff, ((unsigned long)rand() << 16 ^ (unsigned long)rand()) & 0xffffffffffffUL);
        if (!write_out(target + name, csv.data(), nbytes, gz)) { fprintf(stderr, "reflexiv: cannot write %s\n", target.c_str()); rfx_destroy(c); return 1; }
    } else {
        if (rfx_assemble(c) != RFX_OK) return fail(c, "rfx_assemble");
        uint64_t n = 0, total = 0;
        if (rfx_contigs_size(c, &n, &total) != RFX_OK) return fail(c, "rfx_contigs_size");
        std::vector<char> bases(total ? total : 1);
        std::vector<uint64_t> offs(n + 1);
        std::vector<int32_t> left(n ? n : 1), right(n ? n : 1);
        if (rfx_contigs_copy(c, bases.data(), offs.data(), left.data(), right.data()) != RFX_OK) return fail(c, "rfx_contigs_copy");
        std::string out;
        out.reserve(total + total / 100 + 64 * n + 16);
        for (uint64_t i = 0; i < n; i++) {  // DSKmerToContig + changeLine + TagRowContigID, ReflexivDSMain.java:743-794, 717-725
            const uint64_t len = offs[i + 1] - offs[i];
            char head[96];
            snprintf(head, sizeof(head), ">Contig-%llu-(%d,%d)-%llu\n", (unsigned long long)len, left[i], right[i], (unsigned long long)i);
            out += head;
            for (uint64_t j = 0; j < len; j += 100) {
                out.append(bases.data() + offs[i] + j, (size_t)std::min<uint64_t>(100, len - j));
                out.push_back('\n');
            }
        }
        mkdir(target.c_str(), 0755);
        if (!write_out(target + "/part-00000", out.data(), out.size(), gz && from_kmer)) { fprintf(stderr, "reflexiv: cannot write %s\n", target.c_str()); rfx_destroy(c); return 1; }
    }
    FILE* ok = fopen((target + "/_SUCCESS").c_str(), "wb");
    if (ok) fclose(ok);
    rfx_stats_t s;
    rfx_stats(c, &s);
    char msg[512];
    snprintf(msg, sizeof(msg), "done: %llu reads, %llu k-mers, %llu distinct, %llu rows, %llu contigs; GPU ms parse %.2f partition %.2f count %.2f graph %.2f extend %.2f contigs %.2f",
             (unsigned long long)s.n_reads, (unsigned long long)s.n_instances, (unsigned long long)s.n_distinct, (unsigned long long)s.n_rows,
             (unsigned long long)s.n_contigs, s.ms_parse, s.ms_partition, s.ms_count, s.ms_graph, s.ms_extend, s.ms_contigs);
    info(msg);
    rfx_destroy(c);
    return 0;
}
