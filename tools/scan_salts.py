#!/usr/bin/env python
"""How often does the on-image channel itself lose a message?  The reference draws a random salt per embed (S:927-929),
so header and ciphertext bits differ from run to run and a small cover occasionally loses a Rep-3 header bit or a Rep-7
payload bit (upstream behaves identically: same pixels, same raw bits).  Prints, for the covers of
tests/test_pipeline.py and salts s = bytes([k]*16), which round trips succeed -- the test pins salts that do.

    python tools/scan_salts.py [nsalts]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import steganosaurus_b200 as sb  # noqa: E402
from steganosaurus_b200 import host, synth  # noqa: E402

PASS = b"correct horse battery staple"
SPEC = [(256, 256, b"the eagle has landed"), (256, 256, b"second message, same length"[:20]), (256, 256, b"a longer secret " * 4),
        (512, 256, b"second shape"), (512, 256, b"same shape, other length")]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    with sb.Context(0) as ctx:
        for i, (w, h, s) in enumerate(SPEC):
            cover = synth.gen_cover(w, h, 20 + i)
            ok = []
            for k in range(n):
                stego, _ = host.embed_image(ctx, cover, s, PASS, pbkdf2_iter=1000, salt=bytes([k] * 16))
                try:
                    ok.append(int(host.extract_image(ctx, stego, PASS, pbkdf2_iter=1000) == s))
                except host.ExtractError:
                    ok.append(0)
            print(json.dumps({"cover": f"{w}x{h} seed {20 + i}", "secret_len": len(s), "ok_by_salt": ok, "rate": sum(ok) / n}), flush=True)


if __name__ == "__main__":
    main()
