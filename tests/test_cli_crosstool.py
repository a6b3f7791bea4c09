"""Cross-tool drop-in tests (GPU): the new `turtlefft` CLI (CUDA hot path) against the reference CLI
compiled from the unmodified sources (oracle/_ref/turtlefft).  Reference-embedded images must
extract with the new tool and vice versa, with the same messages and exit codes.

Determinism.  The on-image channel is lossy: the reference's OWN round trip loses a Rep-3 header bit on ~7 % of
random salts at 2048^2 (raw BER 0.3 %).  So nothing here asserts "the message comes back" under a random salt.
The asserted property is AGREEMENT: on the same stego image both tools (or our path and the oracle) produce the
same header bytes, payload bytes, message / error and exit code.  Our embeds pin the salt (TFFT_TEST_SALT_HEX for
the CLI, salt= for the Python driver) to values checked on the reference's CPU path (stego pixels identical with
the oracle's), including one salt on which the reference itself fails."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
from steganosaurus_b200 import host, synth
import steganosaurus_b200 as sb

pytestmark = pytest.mark.gpu

OURS = os.path.join(os.path.dirname(os.path.abspath(sb.__file__)), "turtlefft")
REF = O.REF_CLI
PASS = "correct horse battery staple"
MSG = "the eagle has landed"


def run(exe, *args, timeout=300, salt=None):
    env = dict(os.environ)
    env.pop("TFFT_TEST_SALT_HEX", None)
    if salt is not None:
        env["TFFT_TEST_SALT_HEX"] = salt.hex()
    p = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout, env=env)
    return p.returncode, p.stdout, p.stderr


@pytest.fixture(scope="module")
def tools():
    if not os.path.exists(REF):
        pytest.skip("reference CLI not shipped (oracle/_ref/turtlefft)")
    assert os.path.exists(OURS), "steganosaurus_b200/turtlefft not built"
    return OURS, REF


@pytest.fixture(scope="module")
def cover512(tmp_path_factory):
    p = str(tmp_path_factory.mktemp("cli") / "host512.png")
    host.png_save(p, synth.gen_cover(512, 512, 7))
    return p


# (flags, pinned salt, recovery asserted).  Salts with recover=True were checked on the reference's CPU path for this
# cover: at most one raw header error, message recovered (tools/scan_salts.py documents the scan).  Small --alpha
# (raw BER ~2 % on this cover) and --jitter cases assert agreement only.
C1_CASES = [([], bytes([3]) * 16, True), (["--center", "1"], bytes([7]) * 16, True),
            (["--alpha", "0.18", "--density", "0.5"], bytes([1]) * 16, False),
            (["--rmin", "0.1", "--rmax", "0.3"], bytes([0]) * 16, True),
            (["--jitter", "0.05"], bytes([2]) * 16, False), (["--alpha", "0.22", "--jitter", "0.05"], bytes([4]) * 16, False)]


@pytest.mark.parametrize("flags,salt,recover", C1_CASES, ids=[" ".join(c[0]) or "defaults" for c in C1_CASES])
def test_c1_both_tools_agree_on_every_stego(tools, cover512, tmp_path, flags, salt, recover):
    ours, ref = tools
    common = ["--pass", PASS, "--pbkdf2_iter", "1000", *flags]
    s_ours, s_ref = str(tmp_path / "ours.png"), str(tmp_path / "ref.png")
    rc, out, err = run(ours, "embed", "--in", cover512, "--out", s_ours, "--secret", MSG, *common, salt=salt)
    assert rc == 0, err
    assert out.strip() == f"Embedded 2928 bits into {s_ours} (payload 20 bytes, ver=2, salt/nonce in header)"
    rc, out, err = run(ref, "embed", "--in", cover512, "--out", s_ref, "--secret", MSG, *common)  # random salt upstream (S:927-929)
    assert rc == 0, err
    # Drop-in property: on the SAME stego file both tools behave identically -- message, error text and exit code.
    for stego in (s_ours, s_ref):
        res = [run(tool, "extract", "--in", stego, *common) for tool in (ours, ref)]
        assert res[0] == res[1], (stego, res)
    if recover:  # the pinned salt is one the reference's own CPU path recovers
        assert run(ours, "extract", "--in", s_ours, *common)[:2] == (0, MSG + "\n")


def test_same_salt_gives_identical_stego_pixels(tools, cover512, tmp_path):
    """With the salt pinned the new CLI's stego image equals the oracle's embed of the same frame."""
    ours, _ = tools
    s = str(tmp_path / "fixed.png")
    salt = bytes(range(16))
    rc, _, err = run(ours, "embed", "--in", cover512, "--out", s, "--secret", MSG, "--pass", "pw", "--pbkdf2_iter", "1000", salt=salt)
    assert rc == 0, err
    r = O.ref()
    bits, _ = r.frame_bits(b"pw", salt, 1000, MSG.encode())
    bins, _, _ = r.walk(b"pw", 512, 512, bits.size)
    want = r.embed(host.png_load(cover512), bins, bits)["stego"]
    got = host.png_load(s)
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.9999


def test_failure_messages_match_reference(tools, cover512, tmp_path):
    ours, ref = tools
    s = str(tmp_path / "s.png")
    assert run(ours, "embed", "--in", cover512, "--out", s, "--secret", MSG, "--pass", PASS, "--pbkdf2_iter", "1000", salt=bytes([3]) * 16)[0] == 0
    for tool in (ours, ref):
        rc, out, err = run(tool, "extract", "--in", s, "--pass", "wrong", "--pbkdf2_iter", "1000")
        assert (rc, out, err) == (1, "", "Magic not found.\n"), tool                                    # S:1237
        rc, out, err = run(tool, "extract", "--in", s, "--pass", PASS, "--pbkdf2_iter", "1001")
        assert (rc, out, err) == (1, "", "Auth failed (wrong pass or data corrupted).\n"), tool       # S:1308
        rc, out, err = run(tool, "extract", "--in", s, "--pass", PASS, "--pbkdf2_iter", "1000", "--density", "0.5")
        assert (rc, err) == (1, "Magic not found.\n"), tool
        rc, out, err = run(tool, "extract", "--in", str(tmp_path / "nope.png"), "--pass", PASS)
        assert rc == 1 and err == f"Failed to load {tmp_path / 'nope.png'}\n"
        rc, out, err = run(tool, "embed", "--in", cover512, "--bogus", "1")
        assert rc == 1 and err.startswith("Unknown arg: --bogus\n")


def test_key_option_both_ways(tools, cover512, tmp_path):
    """--key (S:576-591, S:1020-1040): a raw master key instead of a passphrase; path_key = SHA256(master_key).
    Both tools agree on each other's stego images; a different key finds no magic."""
    ours, ref = tools
    key = "AAECAwQFBgcICQoLDA0ODxAREhMUFRYXGBkaGxwdHh8="
    other = "AQECAwQFBgcICQoLDA0ODxAREhMUFRYXGBkaGxwdHh8="
    s_ours, s_ref = str(tmp_path / "ours.png"), str(tmp_path / "ref.png")
    assert run(ours, "embed", "--in", cover512, "--out", s_ours, "--secret", MSG, "--key", key, salt=bytes([5]) * 16)[0] == 0
    assert run(ref, "embed", "--in", cover512, "--out", s_ref, "--secret", MSG, "--key", key)[0] == 0
    for stego in (s_ours, s_ref):
        res = [run(tool, "extract", "--in", stego, "--key", key) for tool in (ours, ref)]
        assert res[0] == res[1], (stego, res)
        res = [run(tool, "extract", "--in", stego, "--key", other) for tool in (ours, ref)]
        assert res[0] == res[1] == (1, "", "Magic not found.\n"), (stego, res)
    for tool in (ours, ref):
        rc, out, err = run(tool, "extract", "--in", s_ours, "--key", "@@@")
        assert rc == 1 and err.endswith("Failed to decode/unwrap key from --key argument\n"), (tool, err)


def test_capacity_message_matches_reference(tools, tmp_path):
    ours, ref = tools
    c = str(tmp_path / "host256.png")
    host.png_save(c, synth.gen_cover(256, 256, 5))
    secret = "x" * 400
    outs = []
    for tool in (ours, ref):
        rc, out, err = run(tool, "embed", "--in", c, "--out", str(tmp_path / "o.png"), "--secret", secret, "--pass", PASS,
                           "--pbkdf2_iter", "1000")
        assert rc == 1 and out == ""
        outs.append(err)
    assert outs[0] == outs[1] == "Message too large. Need 24208 bits (after ECC), capacity ~15288 bits.\n"   # S:1010


def test_non_pow2_fails_like_reference(tools, tmp_path):
    """on padded sizes the crop destroys the signal and the reference's own extract fails."""
    ours, ref = tools
    c = str(tmp_path / "c.png")
    host.png_save(c, synth.gen_cover(640, 360, 3))
    s = str(tmp_path / "s.png")
    assert run(ours, "embed", "--in", c, "--out", s, "--secret", MSG, "--pass", PASS, "--pbkdf2_iter", "1000", salt=bytes([1]) * 16)[0] == 0
    res = [run(tool, "extract", "--in", s, "--pass", PASS, "--pbkdf2_iter", "1000") for tool in (ours, ref)]
    assert res[0] == res[1]
    assert (res[0][0], res[0][2]) == (1, "Magic not found.\n")


def test_garbage_length_fails_cleanly(ctx):
    """A decoded header whose length cannot fit the annulus is refused before anything is allocated or walked
    (clen is noise- or attacker-controlled; the reference walks forever here)."""
    cover = synth.gen_cover(256, 256, 5)
    salt = bytes(16)
    hdr = b"FTTG\x02\x00" + salt + bytes(12) + (0xFFFFFFF0).to_bytes(4, "big")
    bits = np.repeat(np.unpackbits(np.frombuffer(hdr, np.uint8)), 3).astype(np.uint8)
    bins = host.cached_walk(PASS.encode(), 256, 256, 912, 0.05, 0.45, 0.7)
    stego, _, _ = ctx.embed_batch(cover[None], bins, bits[None])
    with pytest.raises(host.ExtractError, match="Payload truncated|Magic not found"):
        host.extract_image(ctx, stego[0], PASS.encode(), pbkdf2_iter=1000)


# ---- C2, pow2 variant: 2048x2048, 8192-byte payload, through the Python mirror of do_embed / do_extract ------------------
def _oracle_extract_outcome(r, stego, pw, iters):
    """do_extract (S:1112-1312) on the reference's own hot path (oracle/_ref) -> (header bytes, payload bytes | None, outcome)."""
    H, W, _ = stego.shape
    hb = host.cached_walk(pw, H, W, 912, 0.05, 0.45, 0.7)
    _, raw = r.extract(stego, hb, 1)
    hdr = r.rep_decode(raw, 3).tobytes()
    rc, clen, _, _ = host.parse_header(hdr)
    if rc:
        return hdr, None, "Magic not found." if rc == 1 else "Unsupported version"
    nb = 912 + 56 * (clen + 16)
    if nb > 3 * H * W // 2:
        return hdr, None, "Payload truncated after ECC decode."
    allb = host.cached_walk(pw, H, W, nb, 0.05, 0.45, 0.7)
    _, raw = r.extract(stego, allb[912:], 1)
    pay = r.rep_decode(raw, 7).tobytes()
    ok, pt = host.open_payload(pw, iters, hdr, pay, clen)
    return hdr, pay, pt if ok else "Auth failed (wrong pass or data corrupted)."


def _our_extract_outcome(ctx, stego, pw, iters):
    H, W, _ = stego.shape
    ctx.forward_batch(stego[None])
    hdr = ctx.read_bits(host.cached_walk(pw, H, W, 912, 0.05, 0.45, 0.7), 3, want_raw=False)[0][0].tobytes()
    rc, clen, _, _ = host.parse_header(hdr)
    pay = None
    if rc == 0 and 912 + 56 * (clen + 16) <= 3 * H * W // 2:
        allb = host.cached_walk(pw, H, W, 912 + 56 * (clen + 16), 0.05, 0.45, 0.7)
        pay = ctx.read_bits(allb[912:], 7, want_raw=False)[0][0].tobytes()
    try:
        out = host.extract_image(ctx, stego, pw, pbkdf2_iter=iters)
    except host.ExtractError as e:
        out = str(e)
    return hdr, pay, out


# rng(1002): the reference's CPU path recovers the message (1 raw header error); rng(1010): the reference ITSELF loses a
# Rep-3 header bit (4 raw header errors) -- both tools must fail the same way there
@pytest.mark.parametrize("seed,recovers", [(1002, True), (1010, False)])
def test_c2_pow2_8kb_payload_python_driver(ctx, seed, recovers):
    if not O.have_ref():
        pytest.skip("oracle/_ref not shipped")
    r = O.ref()
    pw = PASS.encode()
    secret = bytes(np.random.default_rng(1).integers(32, 127, 8192, dtype=np.uint8))
    cover = synth.gen_cover(2048, 2048, 1)  # the reference-style fixture (raw BER ~0.3 %)
    salt = np.random.default_rng(seed).bytes(16)
    stego, nbits = host.embed_image(ctx, cover, secret, pw, pbkdf2_iter=1000, salt=salt)
    assert nbits == 460560
    # embed parity: the oracle's stego of the same frame
    bits, _ = r.frame_bits(pw, salt, 1000, secret)
    bins = host.cached_walk(pw, 2048, 2048, nbits, 0.05, 0.45, 0.7)
    want = r.embed(cover, bins, bits)["stego"]
    d = np.abs(stego.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.9999
    # extract parity on the reference's stego: same header bytes, payload bytes and outcome
    ours, theirs = _our_extract_outcome(ctx, want, pw, 1000), _oracle_extract_outcome(r, want, pw, 1000)
    assert ours == theirs
    assert (theirs[2] == secret) == recovers, theirs[2][:60]
    if recovers:
        assert host.extract_image(ctx, stego, pw, pbkdf2_iter=1000) == secret
        with pytest.raises(host.ExtractError, match="Magic not found"):
            host.extract_image(ctx, stego, b"wrong", pbkdf2_iter=1000)
        with pytest.raises(host.ExtractError, match="Auth failed"):
            host.extract_image(ctx, stego, pw, pbkdf2_iter=999)
