"""An independent, strict check of the files Camera::ToFITS writes (the reference goes through EleFits / cfitsio,
src/Camera.cpp:252-307; neither library is in this image, so the writer here is hand-made).

Nothing in this file shares code with the writer (birp_cpp_b200/src/Camera.cpp) or with the small reader the other
tests use (tests/mirror.py: read_fits).  It checks
  * the container rules of the FITS standard (4.0, sections 3 and 4): 2880-byte blocks, 80-character cards of
    printable ASCII, fixed-format values (value indicator in columns 9-10, logical T in column 30, integers and reals
    right-justified in columns 11-30, strings quoted from column 11), the mandatory keywords in their mandatory order,
    END followed by blanks only, big-endian IEEE data, zero padding;
  * every card, byte for byte, against what cfitsio itself emits for the calls EleFits makes on behalf of
    src/Camera.cpp:252-307 -- re-derived here from cfitsio's documented formatting rules (putkey.c: ffpkyl / ffpkyj /
    ffpkye with 7 significant digits and the added decimal point / ffpkys / ffpcom; EleFits puts a record's unit in
    square brackets in front of its comment).
No GPU is needed: the image is filled by hand."""
import re
import struct

import numpy as np
import pytest

import mirror as M

BLOCK, CARD = 2880, 80


def cfitsio_float(x: float) -> str:
    """value string of a TFLOAT keyword (ffr2e with decim = -7)"""
    v = "%.7G" % float(np.float32(x))
    if "." not in v and "E" in v:
        v = "%.1E" % float(np.float32(x))
    if "." not in v and "E" not in v:
        v += "."
    return v


def cfitsio_card(key: str, value: str, comment: str) -> str:
    """ffmkky for a fixed-format value that is not a string: keyword, '= ', the value right-justified to column 30,
    ' / ' and the comment, cut to 80 characters"""
    card = "%-8s= %20s" % (key, value)
    if comment:
        card += " / " + comment
    return card[:CARD].ljust(CARD)


def cfitsio_string_card(key: str, value: str, comment: str) -> str:
    """ffpkys: the string padded to at least 8 characters inside the quotes, the comment aligned as for numbers"""
    quoted = "'%-8s'" % value
    card = "%-8s= %-20s" % (key, quoted)
    if comment:
        card += " / " + comment
    return card[:CARD].ljust(CARD)


def expected_cards(w, h, px, fov_h, fov_w, pos, aim):
    f32 = np.float32
    cards = [
        cfitsio_card("SIMPLE", "T", "file does conform to FITS standard"),
        cfitsio_card("BITPIX", "-32", "number of bits per data pixel"),
        cfitsio_card("NAXIS", "2", "number of data axes"),
        cfitsio_card("NAXIS1", str(w), "length of data axis 1"),
        cfitsio_card("NAXIS2", str(h), "length of data axis 2"),
        cfitsio_card("EXTEND", "T", "FITS dataset may contain extensions"),
        ("COMMENT   FITS (Flexible Image Transport System) format is defined in 'Astronomy").ljust(CARD),
        ("COMMENT   and Astrophysics', volume 376, page 359; bibcode: 2001A&A...376..359H").ljust(CARD),
        cfitsio_card("CRVAL1", cfitsio_float(-(f32(fov_w) / f32(2))), "[deg]"),
        cfitsio_card("CDELT1", cfitsio_float(px), "[deg]"),
        cfitsio_card("CRVAL2", cfitsio_float(-(f32(fov_h) / f32(2))), "[deg]"),
        cfitsio_card("CDELT2", cfitsio_float(px), "[deg]"),
    ]
    for name, v in zip(("POS_X", "POS_Y", "POS_Z", "AIM_X", "AIM_Y", "AIM_Z"), list(pos) + list(aim)):
        cards.append(cfitsio_card(name, cfitsio_float(v), "[float]"))
    cards.append(cfitsio_string_card("RHO", "0", "[t] comment"))
    cards.append("END".ljust(CARD))
    return cards


def strict_parse(raw: bytes):
    """the file as the standard defines it: returns (ordered list of (keyword, value text, comment), data array)"""
    assert len(raw) % BLOCK == 0, "a FITS file is a whole number of 2880-byte blocks"
    cards, off, end = [], 0, None
    while end is None:
        assert off + BLOCK <= len(raw), "no END card"
        for c in range(0, BLOCK, CARD):
            card = raw[off + c:off + c + CARD]
            assert all(0x20 <= b <= 0x7E for b in card), "header cards are printable ASCII"
            text = card.decode("ascii")
            key = text[:8]
            assert re.fullmatch(r"[A-Z0-9_-]* *", key), f"bad keyword field {key!r}"
            if key.rstrip() == "END":
                assert text[8:].strip() == "", "END card must be blank after the keyword"
                end = off + c
                break
            cards.append(text)
        off += BLOCK
    # after END: blanks to the end of the header block
    assert raw[end + CARD:off].strip(b" ") == b"", "the header is padded with blanks"
    parsed = []
    for text in cards:
        key = text[:8].rstrip()
        if key in ("COMMENT", "HISTORY", ""):
            assert text[8:10] != "= ", "commentary keywords carry no value indicator"
            parsed.append((key, None, text[8:].rstrip()))
            continue
        assert text[8:10] == "= ", f"{key}: value indicator must be '= ' in columns 9-10"
        body = text[10:]
        if body.lstrip().startswith("'"):
            assert body[0] == "'", f"{key}: a string value starts in column 11"
            m = re.match(r"'((?:[^']|'')*)'", body)
            assert m, f"{key}: unterminated string"
            assert m.end() >= 10, f"{key}: the closing quote is at or after column 20"
            value, rest = "'" + m.group(1) + "'", body[m.end():]
        else:
            field, rest = body[:20], body[20:]
            value = field.strip()
            assert field == value.rjust(20), f"{key}: fixed-format values are right-justified in columns 11-30"
            if value in ("T", "F"):
                assert text[29] == value, f"{key}: a logical value sits in column 30"
            else:
                assert re.fullmatch(r"[+-]?(\d+\.?\d*|\.\d+)([ED][+-]?\d+)?", value), f"{key}: bad number {value!r}"
        comment = None
        if rest.strip():
            assert rest.lstrip().startswith("/"), f"{key}: text after the value must be a comment"
            comment = rest.lstrip()[1:].strip()
        parsed.append((key, value, comment))
    keys = [k for k, _, _ in parsed]
    assert keys[:3] == ["SIMPLE", "BITPIX", "NAXIS"], "mandatory keywords come first, in order"
    vals = {k: v for k, v, _ in parsed if v is not None}
    assert vals["SIMPLE"] == "T"
    naxis = int(vals["NAXIS"])
    assert keys[3:3 + naxis] == [f"NAXIS{i + 1}" for i in range(naxis)], "NAXISn follow NAXIS immediately"
    assert len(set(k for k in keys if k not in ("COMMENT", "HISTORY", ""))) == len([k for k in keys if k not in ("COMMENT", "HISTORY", "")])
    bitpix = int(vals["BITPIX"])
    assert bitpix in (8, 16, 32, 64, -32, -64)
    shape = [int(vals[f"NAXIS{i + 1}"]) for i in range(naxis)]
    nbytes = abs(bitpix) // 8 * int(np.prod(shape))
    padded = (nbytes + BLOCK - 1) // BLOCK * BLOCK
    assert len(raw) == off + padded, "data are padded to a whole block and nothing follows"
    assert raw[off + nbytes:].strip(b"\0") == b"", "the data padding is zeros"
    assert bitpix == -32
    data = np.array(struct.unpack(">%df" % (nbytes // 4), raw[off:off + nbytes]), dtype=np.float32)
    return parsed, cards, data.reshape(shape[::-1])


@pytest.mark.parametrize("px,fov_h,fov_w,pos,aim", [
    (0.25, 16.0, 27.0, (5.91813, 6.74116, 17.7067), (7.81768, 0.0, 0.0)),
    (0.25, 36.0, 36.0, (30.0, -25.0, 28.0), (7.81768, 3.5, -2.0)),
    (1.0, 7.0, 5.0, (1e-8, -123456.789, 3.0e10), (0.0, -0.0, 1.0)),
    (0.5, 1.0, 720.5, (100.0, 2.5e-5, -7.0), (12345678.0, 1.0, 0.1)),   # 1441 x 2 pixels: a short last block
])
def test_tofits_is_a_strictly_valid_fits_file_with_cfitsio_cards(tmp_path, px, fov_h, fov_w, pos, aim):
    w, h = M.fits_of_image(None, None, pos, aim, px, fov_h, fov_w)
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.normal(size=(h, w)).astype(np.float32)
    img[0, 0] = 0.0
    img[-1, -1] = np.float32(3.4e38)
    if img.size > 4:
        img[0, 1] = -0.0
        img[0, 2] = np.float32(1e-45)
    M.fits_of_image(tmp_path / "img", img, pos, aim, px, fov_h, fov_w)
    raw = (tmp_path / "img.fits").read_bytes()
    parsed, cards, data = strict_parse(raw)
    assert data.shape == (h, w)
    assert np.array_equal(data.view(np.uint32), img.view(np.uint32)), "pixels, bit for bit, row j = image y, x fastest"
    want = expected_cards(w, h, px, fov_h, fov_w, pos, aim)
    got = cards + ["END".ljust(CARD)]
    assert len(got) == len(want)
    for g, e in zip(got, want):
        assert g == e, f"card differs from what cfitsio writes:\n  got  {g!r}\n  want {e!r}"
    # the header values read back as the numbers that went in (7 significant digits)
    vals = {k: v for k, v, _ in parsed if v is not None}
    for name, v in zip(("POS_X", "POS_Y", "POS_Z", "AIM_X", "AIM_Y", "AIM_Z"), list(pos) + list(aim)):
        assert float(vals[name].replace("D", "E")) == pytest.approx(float(np.float32(v)), rel=1e-6, abs=0)
    assert vals["RHO"].strip("'").strip() == "0"


def test_tofits_overwrites_an_existing_file(tmp_path):
    """the reference deletes and rewrites an existing file (src/Camera.cpp:287-301)"""
    pos, aim = (5.0, 6.0, 7.0), (1.0, 0.0, 0.0)
    w, h = M.fits_of_image(None, None, pos, aim, 1.0, 3.0, 4.0)
    M.fits_of_image(tmp_path / "a", np.ones((h, w), np.float32), pos, aim, 1.0, 3.0, 4.0)
    M.fits_of_image(tmp_path / "a", np.full((h, w), 2.0, np.float32), pos, aim, 1.0, 3.0, 4.0)
    _, _, data = strict_parse((tmp_path / "a.fits").read_bytes())
    assert (data == 2.0).all()
