"""Derives the RGB values of the reference's default copper spectra (MetalMaterial's eta / k)
and writes tests/golden/copper_rgb.json.

Run in the build container only: it READS the tables where they lie in /root/reference
(material/metal.rs COPPER_*; spectrum.rs CIE_*), restates RGBSpectrum::from_sampled
(spectrum.rs:2701-2727: interpolate at every CIE wavelength, integrate against the matching
curves, scale by (lambda_last - lambda_first) / (CIE_Y_INTEGRAL * N), XYZ -> RGB) and stores only
the six resulting numbers — no table is copied into this repository."""
import json
import re
import sys
from pathlib import Path

REF = Path("/root/reference/src")


def table(text, name):
    m = re.search(r"pub const %s\s*:\s*\[f64;[^\]]*\]\s*=\s*\[(.*?)\];" % name, text, re.S)
    body = re.sub(r"//.*", "", m.group(1))
    return [float(x.strip().replace("_", "").replace("f64", "")) for x in body.replace("\n", " ").split(",") if x.strip()]


def interpolate(lam, vals, l):  # spectrum.rs:2092-2105
    n = len(lam)
    if l <= lam[0]:
        return vals[0]
    if l >= lam[n - 1]:
        return vals[n - 1]
    off = max(i for i in range(n) if lam[i] <= l)
    off = min(off, n - 2)
    t = (l - lam[off]) / (lam[off + 1] - lam[off])
    return vals[off] * (1.0 - t) + vals[off + 1] * t


def from_sampled(lam, v, cie_l, cx, cy, cz, y_int):
    xyz = [0.0, 0.0, 0.0]
    for i in range(len(cie_l)):
        val = interpolate(lam, v, cie_l[i])
        xyz[0] += val * cx[i]
        xyz[1] += val * cy[i]
        xyz[2] += val * cz[i]
    scale = (cie_l[-1] - cie_l[0]) / (y_int * len(cie_l))
    xyz = [c * scale for c in xyz]
    return [3.240479 * xyz[0] - 1.537150 * xyz[1] - 0.498535 * xyz[2],
            -0.969256 * xyz[0] + 1.875991 * xyz[1] + 0.041556 * xyz[2],
            0.055648 * xyz[0] - 0.204043 * xyz[1] + 1.057311 * xyz[2]]


if __name__ == "__main__":
    if not REF.exists():
        sys.exit("needs /root/reference (build container only)")
    metal = (REF / "material" / "metal.rs").read_text()
    spec = (REF / "spectrum.rs").read_text()
    lam = table(metal, "COPPER_WAVELENGTHS")
    n_s, k_s = table(metal, "COPPER_N_SAMPLES"), table(metal, "COPPER_K_SAMPLES")
    cie = [table(spec, "CIE_LAMBDA"), table(spec, "CIE_X"), table(spec, "CIE_Y"), table(spec, "CIE_Z")]
    y_int = float(re.search(r"CIE_Y_INTEGRAL: f64 = ([0-9.]+)", spec).group(1))
    assert len(lam) == len(n_s) == len(k_s) == 56 and all(len(c) == 471 for c in cie)
    out = {"eta": from_sampled(lam, n_s, *cie, y_int), "k": from_sampled(lam, k_s, *cie, y_int),
           "source": "material/metal.rs COPPER_* through spectrum.rs:2701-2727 (RGBSpectrum::from_sampled)"}
    (Path(__file__).resolve().parent / "copper_rgb.json").write_text(json.dumps(out, indent=1) + "\n")
    print(out)
