// Test harness for the command-line driver's host-side I/O (no GPU needed): the input Source (plain / gzip / blocked gzip,
// whole batches into a caller's buffer) and the blocked-gzip writer, driven exactly as faqcs_cli.cpp drives them.
//   host_io_harness read  IN OUT BUFFER_KIB     fill buffers of that size until EOF, append what arrives to OUT
//   host_io_harness write IN OUT                IN -> BGZF in two write_bgzf calls + the end-of-file member
#define main faqcs_cli_main
#include "../faqcs_b200/host/faqcs_cli.cpp"
#undef main

int main(int argc, char **argv)
{
    if (argc < 4) return 64;
    const std::string mode = argv[1];
    try {
        if (mode == "read") {
            Source s;
            if (!s.open(argv[2])) return 2;
            const bool was_pgzip = s.pgzip, was_bgzf = s.bgzf;
            const size_t cap = (size_t)atoi(argv[4]) << 10;
            std::vector<uint8_t> buf(cap);
            FILE *f = fopen(argv[3], "wb");
            size_t fills = 0, total_lines = 0;
            while (!s.eof) {
                size_t lines = 0;
                const size_t n = s.fill(buf.data(), 0, cap, &lines);
                if (count_newlines(buf.data(), n) != lines) { fprintf(stderr, "newline count mismatch\n"); return 3; }
                if (n == 0 && !s.eof) { fprintf(stderr, "no progress\n"); return 4; }
                fwrite(buf.data(), 1, n, f);
                total_lines += lines;
                ++fills;
            }
            fclose(f);
            printf("fills %zu lines %zu mode %s\n", fills, total_lines, was_pgzip ? "pgzip" : was_bgzf ? "bgzf" : s.gz ? "zlib" : "plain");
            return 0;
        }
        if (mode == "write") {
            FILE *f = fopen(argv[2], "rb");
            if (!f) return 2;
            fseek(f, 0, SEEK_END);
            const size_t n = (size_t)ftell(f);
            fseek(f, 0, SEEK_SET);
            std::vector<uint8_t> b(n);
            if (n && fread(b.data(), 1, n, f) != n) return 2;
            fclose(f);
            const int fd = open_out(argv[3], "test output");
            const size_t cut = n / 3;
            write_bgzf(fd, b.data(), cut);
            write_bgzf(fd, b.data() + cut, n - cut);
            close_bgzf(fd);
            ::close(fd);
            return 0;
        }
    } catch (const char *e) {
        fprintf(stderr, "error: %s\n", e);
        return 5;
    }
    return 64;
}
