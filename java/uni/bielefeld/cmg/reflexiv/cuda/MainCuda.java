/*
 * MainCuda -- the Java driver north_star asks for: Reflexiv's own option parsing (util/Parameter.java,
 * util/ParameterOfCounter.java, unchanged) in front of libreflexiv_cuda instead of the Spark stages.
 *
 *   java --enable-native-access=ALL-UNNAMED -cp lib/original-Reflexiv-1.0.jar:commons-cli.jar:. \
 *        uni.bielefeld.cmg.reflexiv.cuda.MainCuda run     -fastq 'example/paired_dat*.fq.gz' -outfile out -kmer 31 -cover 3
 *        ...                                      counter -fastq ... -outfile out -kmer 31 [-gzip]
 *        ...                                      sort    -kmerc 'out/Count_31/part*' -outfile out -kmer 31
 *
 * Mirrors main/Main.java:57-79 and main/MainOfCounter.java:58-80 (parse, then run the pipeline) and writes what
 * pipeline/ReflexivDataFrameCounter.java:222-233, pipeline/ReflexivDSMain.java:331-354 / 706-710 and
 * pipeline/ReflexivDSKmerLeftAndRightSorting.java:226-238 write.  `bin/reflexiv` selects it by replacing the class
 * names at bin/reflexiv:252-259.
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no JDK): source only.  The built and tested equivalents are
 * reflexiv_b200/csrc/reflexiv_main.cpp (C++) and `python -m reflexiv_b200` (Python).
 */
package uni.bielefeld.cmg.reflexiv.cuda;

import java.io.ByteArrayOutputStream;
import java.io.IOException;
import java.io.InputStream;
import java.io.OutputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.channels.FileChannel;
import java.nio.charset.StandardCharsets;
import java.nio.file.DirectoryStream;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.Paths;
import java.nio.file.StandardOpenOption;
import java.util.ArrayList;
import java.util.Collections;
import java.util.List;
import java.util.UUID;
import java.util.zip.GZIPInputStream;
import java.util.zip.GZIPOutputStream;

import uni.bielefeld.cmg.reflexiv.util.DefaultParam;
import uni.bielefeld.cmg.reflexiv.util.InfoDumper;
import uni.bielefeld.cmg.reflexiv.util.Parameter;
import uni.bielefeld.cmg.reflexiv.util.ParameterOfCounter;

public final class MainCuda {
    public static void main(String[] args) throws Exception {
        if (args.length == 0) { System.err.println("usage: MainCuda <run|counter|sort> [options]"); System.exit(1); }
        String cmd = args[0];
        String[] rest = java.util.Arrays.copyOfRange(args, 1, args.length);
        InfoDumper info = new InfoDumper();
        info.readParagraphedMessages("Reflexiv " + cmd + " initiating ... \ninterpreting parameters.");
        info.screenDump();
        DefaultParam param = cmd.equals("counter") ? new ParameterOfCounter(rest).importCommandLine() : new Parameter(rest).importCommandLine();
        int device = Integer.getInteger("reflexiv.cuda.device", 0);
        switch (cmd) {
            case "counter" -> counter(param, device);
            case "run" -> run(param, device);
            case "sort" -> sort(param, device);
            default -> { System.err.println("command outside the GPU path: " + cmd); System.exit(1); }
        }
    }

    /** ReflexivDataFrameCounter.assembly(), :139-236 */
    static void counter(DefaultParam p, int device) throws IOException {
        try (ReflexivCuda gpu = new ReflexivCuda(p, true, device)) {
            pushInputs(gpu, p.inputFqPath);
            gpu.count();
            writeCsvDir(Paths.get(p.outputPath, "Count_" + p.kmerSize), gpu.countsCsv(), p.gzip);
        }
    }

    /** ReflexivDSMain.assembly() :123-357, or assemblyFromKmer() :362-713 when -kmerc is given (Pipelines.java:83-84; a -fastq next to it
     *  is only read by the -stitch branch, Parameter.java:571-575) */
    static void run(DefaultParam p, int device) throws IOException {
        boolean fromKmer = p.inputKmerPath != null;
        Path out = fromKmer ? Paths.get(p.outputPath, "Assemble_" + p.kmerSize) : Paths.get(p.outputPath);
        if (Files.exists(out)) throw new IOException("Output directory " + out + " already exists");  // saveAsTextFile refuses
        try (ReflexivCuda gpu = new ReflexivCuda(p, false, device)) {
            if (fromKmer) loadCountTable(gpu, p, p.minKmerCoverage);
            else { pushInputs(gpu, p.inputFqPath); gpu.count(); }
            if (fromKmer && p.stitch) {                     // ReflexivDSMain.java:585-672: the FASTQ is read a second time
                if (p.inputFqPath == null) throw new IOException("-stitch reads the FASTQ (ReflexivDSMain.java:599): give -fastq next to -kmerc");
                gpu.stitchBegin();
                pushInputs(gpu, p.inputFqPath);
                gpu.stitchFinish();
            } else gpu.assemble();
            StringBuilder sb = new StringBuilder();
            ReflexivCuda.Contig[] contigs = gpu.contigs();
            for (int i = 0; i < contigs.length; i++) {      // DSKmerToContig + changeLine + TagRowContigID, ReflexivDSMain.java:743-794, 717-725
                ReflexivCuda.Contig c = contigs[i];
                sb.append(">Contig-").append(c.bases().length()).append("-(").append(c.left()).append(',').append(c.right()).append(")-").append(i).append('\n');
                for (int j = 0; j < c.bases().length(); j += 100) sb.append(c.bases(), j, Math.min(j + 100, c.bases().length())).append('\n');
            }
            Files.createDirectories(out);
            writeFile(out.resolve("part-00000"), sb.toString().getBytes(StandardCharsets.US_ASCII), p.gzip && fromKmer);
            Files.write(out.resolve("_SUCCESS"), new byte[0]);
        }
    }

    /** ReflexivDSKmerLeftAndRightSorting.assemblyFromKmer(), :105-243 */
    static void sort(DefaultParam p, int device) throws IOException {
        try (ReflexivCuda gpu = new ReflexivCuda(p, false, device)) {
            loadCountTable(gpu, p, Integer.MIN_VALUE);                        // only count <= maxcov applies here, :186-193
            byte[] csv = gpu.sortedCsv(p.minErrorCoverage, p.minRepeatFold, p.kmerListInt[p.kmerListInt.length - 1]);
            writeCsvDir(Paths.get(p.outputPath, "Count_" + p.kmerSize + "_sorted"), csv, p.gzip);
        }
    }

    // ---- input ----
    static List<Path> expand(String pattern) throws IOException {
        Path pat = Paths.get(pattern);
        Path dir = pat.getParent() == null ? Paths.get(".") : pat.getParent();
        List<Path> files = new ArrayList<>();
        try (DirectoryStream<Path> ds = Files.newDirectoryStream(dir, pat.getFileName().toString())) {
            for (Path f : ds) {
                if (Files.isDirectory(f)) {
                    try (DirectoryStream<Path> inner = Files.newDirectoryStream(f)) {
                        for (Path g : inner) {
                            String n = g.getFileName().toString();
                            if (!n.startsWith("_") && !n.startsWith(".")) files.add(g);
                        }
                    }
                } else files.add(f);
            }
        }
        if (files.isEmpty()) throw new IOException("Input path does not exist: " + pattern);
        Collections.sort(files);
        return files;
    }

    static byte[] readAll(Path f) throws IOException {
        try (InputStream in = f.toString().endsWith(".gz") ? new GZIPInputStream(Files.newInputStream(f), 1 << 20) : Files.newInputStream(f)) {
            ByteArrayOutputStream bo = new ByteArrayOutputStream();
            in.transferTo(bo);
            return bo.toByteArray();
        }
    }

    /** one rfx_push_fastq per file: plain files are mapped, .gz files inflated into a native buffer */
    static void pushInputs(ReflexivCuda gpu, String pattern) throws IOException {
        for (Path f : expand(pattern)) {
            try (Arena a = Arena.ofConfined()) {
                if (f.toString().endsWith(".gz")) {
                    byte[] text = readAll(f);
                    if (text.length == 0) continue;
                    boolean nl = text[text.length - 1] == '\n';
                    MemorySegment seg = a.allocate(text.length + (nl ? 0 : 1));
                    MemorySegment.copy(text, 0, seg, java.lang.foreign.ValueLayout.JAVA_BYTE, 0, text.length);
                    if (!nl) seg.set(java.lang.foreign.ValueLayout.JAVA_BYTE, text.length, (byte) '\n');
                    gpu.pushFastq(seg);
                } else {
                    try (FileChannel ch = FileChannel.open(f, StandardOpenOption.READ)) {
                        if (ch.size() == 0) continue;
                        MemorySegment seg = ch.map(FileChannel.MapMode.READ_ONLY, 0, ch.size(), a);
                        if (seg.get(java.lang.foreign.ValueLayout.JAVA_BYTE, ch.size() - 1) == '\n') gpu.pushFastq(seg);
                        else {                                                  // no final newline: a copy that gets one
                            MemorySegment copy = a.allocate(ch.size() + 1);
                            copy.copyFrom(seg);
                            copy.set(java.lang.foreign.ValueLayout.JAVA_BYTE, ch.size(), (byte) '\n');
                            gpu.pushFastq(copy);
                        }
                    }
                }
            }
        }
    }

    /** KmerBinarizer input: `KMER,count` or the legacy `(KMER,count)`; >= 10 digits clamp to 10^9 (ReflexivDSMain.java:3895-3910) */
    static void loadCountTable(ReflexivCuda gpu, DefaultParam p, int minCover) throws IOException {
        int k = p.kmerSize, words = k <= 31 ? 1 : k / 32 + 1, res = k % 32;
        List<long[]> keys = new ArrayList<>();
        List<Integer> counts = new ArrayList<>();
        for (Path f : expand(p.inputKmerPath)) {
            for (String line : new String(readAll(f), StandardCharsets.US_ASCII).split("\n")) {
                if (line.isEmpty()) continue;
                if (line.startsWith("(")) line = line.substring(1, line.length() - 1);
                int comma = line.indexOf(',');
                String num = line.substring(comma + 1);
                int cover = num.length() >= 10 ? 1000000000 : Integer.parseInt(num);
                if (cover < minCover || cover > p.maxKmerCoverage) continue;
                long[] key = new long[words];
                for (int i = 0; i < k; i++) {
                    char ch = line.charAt(i);
                    long v = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3;
                    if (words == 1) key[0] = (key[0] << 2) | v;
                    else {                                                       // 32 bases per word, the last word holds k % 32 right aligned
                        int w = i / 32;
                        key[w] = (key[w] << 2) | v;
                    }
                }
                keys.add(key);
                counts.add(cover);
            }
        }
        long[] flat = new long[keys.size() * words];
        int[] cnt = new int[keys.size()];
        for (int i = 0; i < keys.size(); i++) { System.arraycopy(keys.get(i), 0, flat, i * words, words); cnt[i] = counts.get(i); }
        if (res == 0 && words > 1) throw new IOException("k % 32 == 0 is outside the reference's domain (ReflexivDataFrameCounter64)");
        gpu.loadCounts(flat, cnt);
    }

    // ---- output ----
    static void writeFile(Path f, byte[] data, boolean gzip) throws IOException {
        Path target = gzip ? f.resolveSibling(f.getFileName() + ".gz") : f;
        try (OutputStream raw = Files.newOutputStream(target); OutputStream out = gzip ? new GZIPOutputStream(raw, 1 << 16) : raw) {
            out.write(data);
        }
    }

    /** write().mode(SaveMode.Overwrite).csv(dir): one part file + _SUCCESS */
    static void writeCsvDir(Path dir, byte[] csv, boolean gzip) throws IOException {
        Files.createDirectories(dir);
        try (DirectoryStream<Path> ds = Files.newDirectoryStream(dir)) {
            for (Path old : ds) Files.delete(old);
        }
        writeFile(dir.resolve("part-00000-" + UUID.randomUUID() + "-c000.csv"), csv, gzip);
        Files.write(dir.resolve("_SUCCESS"), new byte[0]);
    }
}
