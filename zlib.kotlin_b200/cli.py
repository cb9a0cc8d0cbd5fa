"""Command-line driver with the reference CLI's commands and messages (nativeMain/.../cli/ZLibCli.kt:11-123), running on
the GPU engine:  python -m zlibk_b200.cli compress|deflate <in> <out> [level]  |  decompress|inflate <in> <out>  |  help

The logging toggles of the reference (`log-on`, `log-off`) are accepted and answered, but there is no logger here."""
import os
import sys

from . import api
from .binding import Z_STREAM_END


def print_help(out):
    out.write("ZLib.kotlin - Pure arithmetic zlib implementation\n")
    out.write("Commands:\n")
    out.write("  compress|deflate <input.txt> <output.zz> [level]  - Compress a file (zlib). Level: 1..9 (default 6)\n")
    out.write("  decompress|inflate <input.zz> <output.txt>        - Decompress a zlib file\n")
    out.write("  log-on                                            - Enable logging (DEBUG off by default)\n")
    out.write("  log-off                                           - Disable logging\n")
    out.write("  help                                              - Show this help\n")


def main(argv=None, out=None):
    args = list(sys.argv[1:] if argv is None else argv)
    out = out or sys.stdout
    if not args or args[0] in ("help", "--help", "-h"):
        print_help(out)
        return 0
    cmd = args[0]
    if cmd in ("compress", "deflate"):
        if len(args) < 3:
            out.write("Usage: %s <input.txt> <output.zz> [level]\n" % cmd)
            return 0
        try:
            level = int(args[3]) if len(args) > 3 else 6
        except ValueError:
            level = 6  # toIntOrNull() ?: 6
        with open(args[1], "rb") as src, open(args[2], "wb") as snk:
            bytes_in = api.DeflateStream.compressZlib(src, snk, level)
        out.write("Compressed %d bytes to %d bytes (level=%d)\n" % (bytes_in, os.path.getsize(args[2]), level))
        return 0
    if cmd in ("decompress", "inflate"):
        if len(args) < 3:
            out.write("Usage: %s <input.zz> <output.txt>\n" % cmd)
            return 0
        with open(args[1], "rb") as src, open(args[2], "wb") as snk:
            result, bytes_out = api.InflateStream.inflateZlib(src, snk)
        if result == Z_STREAM_END:
            out.write("Decompressed %d bytes to %d bytes\n" % (os.path.getsize(args[1]), bytes_out))
        else:
            out.write("Decompression failed: %d\n" % result)
        return 0
    if cmd == "log-on":
        out.write("Logging enabled (DEBUG=off, BITWISE=off). Use environment vars ZLIB_LOG_DEBUG=1 and ZLIB_LOG_BITWISE=1 for more detail.\n")
        return 0
    if cmd == "log-off":
        out.write("Logging disabled\n")
        return 0
    out.write("Unknown command: %s\n" % cmd)
    print_help(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
