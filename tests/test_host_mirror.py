"""The C++ host mirror (hipims_ocl_b200/host): same XML schema, class surface and semantics as the
reference's CModel / CDomainCartesian / CScheme* / CBoundary* for the hot path, above the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

from hipims_ocl_b200 import build as hpbuild
from hipims_ocl_b200 import config as hc
from hipims_ocl_b200 import scenarios as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

XML = """<?xml version="1.0"?>
<!DOCTYPE configuration PUBLIC "HiPIMS Configuration Schema 1.1" "http://www.lukesmith.org.uk/research/namespace/hipims/1.1/"[]>
<configuration>
	<metadata>
		<name>Synthetic pluvial test</name>
		<description>Same schema as test/newcastle-centre.xml of the reference</description>
	</metadata>
	<execution>
		<executor name="OpenCL">
			<parameter name="deviceFilter" value="GPU" />
		</executor>
	</execution>
	<simulation>
		<parameter name="duration" value="{duration}" />
		<parameter name="outputFrequency" value="{outfreq}" />
		<parameter name="floatingPointPrecision" value="{precision}" />
		<domainSet>
			<domain type="cartesian" deviceNumber="1">
				<data sourceDir="topography/" targetDir="output/">
					<dataSource type="constant" value="velocityX" source="0.0" />
					<dataSource type="constant" value="velocityY" source="0.0" />
					<dataSource type="constant" value="depth" source="0.0" />
					<dataSource type="constant" value="manningCoefficient" source="0.030" />
					<dataSource type="raster" value="structure,dem" source="dem.asc" />
					<dataTarget type="raster" value="depth" format="GTiff" target="depth_%t.tif" />
					<dataTarget type="raster" value="velocityX" format="ENVI" target="velX_%t.bil" />
					<dataTarget type="raster" value="fsl" format="AAIGrid" target="fsl_%t.asc" />
					<dataTarget type="raster" value="maxdepth" format="HFA" target="maxdepth_%t.img" />
				</data>
				<scheme name="{scheme}">
					<parameter name="courantNumber" value="0.50" />
					<parameter name="frictionEffects" value="yes" />
					<parameter name="groupSize" value="32x8" />
					<parameter name="queueSize" value="16" />
				</scheme>
				<boundaryConditions sourceDir="boundaries/">
					<domainEdge edge="north" treatment="closed" />
					<!-- rain and losses, like the reference's test -->
					<timeseries type="atmospheric" name="Drainage" value="loss-rate" source="drainage.csv" />
					<timeseries type="atmospheric" name="Rainfall" value="rain-intensity" source="rainfall.csv" />
					<timeseries type="cell" name="Inflow" depthValue="ignore" dischargeValue="total" source="inflow.csv" mapFile="inflow_map.csv" />
				</boundaryConditions>
			</domain>
		</domainSet>
	</simulation>
</configuration>
"""


def write_asc(path, a, cellsize, xll=424520.0, yll=565146.0):
    rows, cols = a.shape
    with open(path, "w") as f:
        f.write("ncols %d\nnrows %d\nxllcorner %r\nyllcorner %r\ncellsize %r\nNODATA_value -9999\n" % (cols, rows, xll, yll, cellsize))
        for r in range(rows - 1, -1, -1):           # north first in the file
            f.write(" ".join(repr(float(v)) for v in a[r]) + "\n")


def read_asc(path):
    with open(path) as f:
        hdr = {}
        for _ in range(6):
            k, v = f.readline().split()
            hdr[k.lower()] = float(v)
        data = np.loadtxt(f)
    return data[::-1], hdr


def read_tiff(path):
    """Minimal reader for what CRasterDataset::writeRaster("GTiff") produces: little-endian classic TIFF or BigTIFF,
    uncompressed strips.  Returns (array, tags)."""
    import struct
    raw = open(path, "rb").read()
    assert raw[:2] == b"II"
    magic = struct.unpack_from("<H", raw, 2)[0]
    big = magic == 43
    assert magic in (42, 43)
    if big:
        assert struct.unpack_from("<HH", raw, 4) == (8, 0)
        off = struct.unpack_from("<Q", raw, 8)[0]
        n = struct.unpack_from("<Q", raw, off)[0]; off += 8; esz = 20
    else:
        off = struct.unpack_from("<I", raw, 4)[0]
        n = struct.unpack_from("<H", raw, off)[0]; off += 2; esz = 12
    sizes = {2: 1, 3: 2, 4: 4, 12: 8, 16: 8}
    fmts = {2: "c", 3: "H", 4: "I", 12: "d", 16: "Q"}
    tags, last = {}, 0
    for i in range(n):
        tag, typ = struct.unpack_from("<HH", raw, off + i * esz)
        assert tag > last; last = tag                                   # ascending, as the TIFF specification requires
        cnt = struct.unpack_from("<Q" if big else "<I", raw, off + i * esz + 4)[0]
        voff = off + i * esz + (12 if big else 8)
        if cnt * sizes[typ] > (8 if big else 4):
            voff = struct.unpack_from("<Q" if big else "<I", raw, voff)[0]
        vals = struct.unpack_from("<%d%s" % (cnt, fmts[typ]), raw, voff)
        tags[tag] = b"".join(vals).rstrip(b"\0").decode() if typ == 2 else list(vals)
    cols, rows = tags[256][0], tags[257][0]
    assert tags[258] == [64] and tags[259] == [1] and tags[277] == [1] and tags[339] == [3] and tags[278] == [1]
    data = np.empty((rows, cols))
    for r, (o, c) in enumerate(zip(tags[273], tags[279])):
        assert c == cols * 8
        data[r] = np.frombuffer(raw, dtype="<f8", count=cols, offset=o)
    return data, tags


def read_envi(path):
    hdr = {}
    for line in open(os.path.splitext(path)[0] + ".hdr").read().splitlines()[1:]:
        k, _, v = line.partition("=")
        hdr[k.strip()] = v.strip()
    assert hdr["data type"] == "5" and hdr["byte order"] == "0" and hdr["interleave"] == "bsq" and hdr["bands"] == "1"
    return np.fromfile(path, dtype="<f8").reshape(int(hdr["lines"]), int(hdr["samples"])), hdr


@pytest.fixture()
def model_dir(tmp_path):
    def make(scheme="Godunov", precision="double", duration=30, outfreq=10, rows=30, cols=40):
        for d in ("topography", "output", "boundaries"):
            os.makedirs(tmp_path / d, exist_ok=True)
        bed = sc.fractal_dem(rows, cols, 5, amplitude=3.0) + 0.00004   # not on the 4-decimal grid: ingestion must round
        write_asc(tmp_path / "topography" / "dem.asc", bed, 2.0)
        (tmp_path / "boundaries" / "rainfall.csv").write_text("Time (s),Rainfall intensity (mm/hr)\n0,70\n3600,70\n7200,0\n10800,0\n\n\n")
        (tmp_path / "boundaries" / "drainage.csv").write_text("Time (s),Drainage losses (mm/hr)\n0,12\n100000000,12\n\n")
        (tmp_path / "boundaries" / "inflow.csv").write_text("t,depth,qx,qy\n0,0,0,0\n10,0,3.0,0\n100000,0,3.0,0\n")
        (tmp_path / "boundaries" / "inflow_map.csv").write_text("x,y\n1,10\n1,11\n1,12\n")
        cfg = tmp_path / "model.xml"
        cfg.write_text(XML.format(scheme=scheme, precision=precision, duration=duration, outfreq=outfreq))
        return str(cfg), bed
    return make


@pytest.fixture(scope="module")
def host():
    hpbuild.build()
    hpbuild.build_host()
    from hipims_ocl_b200 import executor as hx
    hx.prefer_bundled_nccl()          # the strips of the multi-device tests must share PyTorch's NCCL (one libnccl per process)
    lib = C.CDLL(hpbuild.HOST_LIB)
    lib.hph_model_load.restype = C.c_void_p
    lib.hph_model_load.argtypes = [C.c_char_p, C.c_int]
    lib.hph_model_run.argtypes = [C.c_void_p]
    lib.hph_model_destroy.argtypes = [C.c_void_p]
    for name in ("hph_model_states", "hph_model_bed", "hph_model_manning"):
        getattr(lib, name).restype = C.POINTER(C.c_double)
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.hph_error.restype = C.c_char_p
    lib.hph_round.restype = C.c_double
    lib.hph_round.argtypes = [C.c_double, C.c_int]
    lib.hph_derive_output.restype = C.c_double
    lib.hph_derive_output.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_double, C.c_double]
    return lib


def info(lib, h):
    cols, rows, bnd = C.c_ulong(), C.c_ulong(), C.c_uint()
    res, dur, freq = C.c_double(), C.c_double(), C.c_double()
    prec, scheme = C.c_int(), C.c_int()
    lib.hph_model_info(C.c_void_p(h), C.byref(cols), C.byref(rows), C.byref(res), C.byref(dur), C.byref(freq), C.byref(prec), C.byref(scheme), C.byref(bnd))
    return dict(cols=cols.value, rows=rows.value, resolution=res.value, duration=dur.value, output_frequency=freq.value,
                precision=prec.value, scheme=scheme.value, boundaries=bnd.value)


def arrays(lib, h, rows, cols):
    n = rows * cols
    st = np.ctypeslib.as_array(lib.hph_model_states(C.c_void_p(h)), shape=(n * 4,)).reshape(rows, cols, 4).copy()
    bed = np.ctypeslib.as_array(lib.hph_model_bed(C.c_void_p(h)), shape=(n,)).reshape(rows, cols).copy()
    man = np.ctypeslib.as_array(lib.hph_model_manning(C.c_void_p(h)), shape=(n,)).reshape(rows, cols).copy()
    return st, bed, man


def boundaries(lib, h, count):
    out = []
    for i in range(count):
        kind, da, db, rel = C.c_int(), C.c_int(), C.c_int(), C.c_uint()
        interval, length = C.c_double(), C.c_double()
        series = C.POINTER(C.c_double)()
        n = lib.hph_model_boundary(C.c_void_p(h), i, C.byref(kind), C.byref(da), C.byref(db), C.byref(interval), C.byref(length), C.byref(series), C.byref(rel))
        out.append(dict(kind=kind.value, def_a=da.value, def_b=db.value, interval=interval.value, length=length.value,
                        relations=rel.value, series=[series[j] for j in range(max(n, 0))]))
    return out


def test_rounding_matches_reference_util_round(host):
    # src/util.cpp:79-93: half up for positives, negatives always towards -infinity
    assert host.hph_round(1.23455, 4) == 1.2346 and host.hph_round(1.23454, 4) == 1.2345
    assert host.hph_round(-1.23451, 4) == -1.2346 and host.hph_round(-0.00001, 4) == -0.0001
    for v in (3.14159265, -2.000049, 0.00005, 17.99995):
        assert host.hph_round(v, 4) == float(sc.round4(v))


def test_rounding_is_the_reference_function(host):
    """The mirror's Util::round and the scenario generator's round4 against the reference's own Util::round, compiled from
    src/util.cpp where it lies (skipped where /root/reference is absent)."""
    from oracle import build_ref
    if not build_ref.reference_available():
        pytest.skip("reference tree not present")
    ref = C.CDLL(build_ref.build_util())
    ref.ref_round.restype = C.c_double
    ref.ref_round.argtypes = [C.c_double, C.c_ubyte]
    rng = np.random.default_rng(4)
    values = np.concatenate([rng.normal(scale=50.0, size=4000), rng.normal(scale=1e-3, size=1000),
                             np.round(rng.normal(scale=10.0, size=1000), 4) + 0.00005, [0.0, -0.0, 1e-9, -1e-9, 9999.9, -9999.0]])
    for v in values:
        want = ref.ref_round(float(v), 4)
        assert host.hph_round(float(v), 4) == want
        assert float(sc.round4(float(v))) == want
    for places in (0, 1, 2, 6):
        for v in values[:500]:
            assert host.hph_round(float(v), places) == ref.ref_round(float(v), places)


def hfa_nodes(path):
    """Walks the node tree of an ERDAS IMAGINE file: {name: (type, data bytes)} plus the MIF dictionary."""
    import struct
    d = open(path, "rb").read()
    assert d[:16] == b"EHFA_HEADER_TAG\0"
    hdr = struct.unpack_from("<I", d, 16)[0]
    version, free, root, ehl, dictp = struct.unpack_from("<iIIhI", d, hdr)
    assert (version, ehl) == (1, 128)
    nodes = {}

    def walk(o, parent):
        prev = 0
        while o:
            nxt, prv, par, child, data, size = struct.unpack_from("<IIIIIi", d, o)
            name = d[o + 24:o + 88].split(b"\0")[0].decode()
            typ = d[o + 88:o + 120].split(b"\0")[0].decode()
            assert par == parent and prv == prev and data + size <= len(d)
            nodes[name] = (typ, d[data:data + size], data)
            if child:
                walk(child, o)
            prev, o = o, nxt
    walk(root, 0)
    return nodes, d[dictp:d.index(b"\0", dictp)].decode(), d


def check_hfa_structure(path, cols, rows):
    """The written file against the layout GDAL's HFA driver produces (the reference's test DEM is such a file: compared
    node by node where /root/reference is present)."""
    import struct
    nodes, dictionary, raw = hfa_nodes(path)
    assert {"root", "IMGFormatInfo", "Layer_1", "RasterDMS", "Ehfa_Layer", "Eimg_NonInitializedValue", "Map_Info"} <= set(nodes)
    for name, (typ, _, _) in nodes.items():
        assert typ == "root" or ("}" + typ + ",") in dictionary            # every object type is defined in the MIF dictionary
    assert dictionary.endswith(",.")
    assert struct.unpack("<iiHHii", nodes["Layer_1"][1]) == (cols, rows, 1, 10, 64, 64)
    nblocks = -(-cols // 64) * -(-rows // 64)
    dms, dms_at = nodes["RasterDMS"][1], nodes["RasterDMS"][2]
    assert struct.unpack_from("<iiiH", dms) == (nblocks, 4096, nblocks * 4096, 0)
    count, ptr = struct.unpack_from("<II", dms, 14)
    assert count == nblocks and ptr == dms_at + 22 and len(dms) == 22 + 14 * nblocks + 16
    for b in range(nblocks):
        code, off, size, valid, comp = struct.unpack_from("<HIiHH", dms, 22 + 14 * b)
        assert (code, size, valid, comp) == (0, 32768, 1, 0) and off + size <= len(raw)
    typ, ldict_ptr = struct.unpack("<HI", nodes["Ehfa_Layer"][1])
    assert typ == 0 and raw[ldict_ptr:raw.index(b"\0", ldict_ptr)] == b"{4096:ddata,}RasterDMS,."
    n, p, r, c, t, o = struct.unpack_from("<IIiiHH", nodes["Eimg_NonInitializedValue"][1])
    assert (n, p, r, c, t, o) == (1, nodes["Eimg_NonInitializedValue"][2] + 8, 1, 1, 10, 0)
    assert struct.unpack_from("<d", nodes["Eimg_NonInitializedValue"][1], 20)[0] == -9999.0
    sample = "/root/reference/test/newcastle-centre/topography/NewcastleCentreDEM_2m.img"
    if os.path.exists(sample):
        ref_nodes, ref_dict, _ = hfa_nodes(sample)
        for name in ("IMGFormatInfo", "Layer_1", "RasterDMS", "Ehfa_Layer", "Eimg_NonInitializedValue", "Map_Info"):
            assert nodes[name][0] == ref_nodes[name][0]                      # same object types as the GDAL-written file
            typ = nodes[name][0]
            i = ref_dict.find("}" + typ + ",")
            definition = ref_dict[ref_dict.rfind("{", 0, i):i + len(typ) + 2]
            assert definition in dictionary                                  # and the same MIF definition of each
        assert len(nodes["Layer_1"][1]) == len(ref_nodes["Layer_1"][1]) and len(nodes["Ehfa_Layer"][1]) == len(ref_nodes["Ehfa_Layer"][1])
        assert len(nodes["Eimg_NonInitializedValue"][1]) == len(ref_nodes["Eimg_NonInitializedValue"][1])


def test_raster_writers_round_trip(host, tmp_path):
    """GeoTIFF / ENVI / ESRI ASCII written by the GDAL-free CRasterDataset stand-in: values bit-exact, georeferencing
    as src/Datasets/CRasterDataset.cpp:163-176 sets it (top-left origin, negative y resolution, no-data -9999)."""
    host.hph_write_raster.argtypes = [C.c_char_p, C.c_char_p, C.c_ulong, C.c_ulong, C.c_double, C.c_double, C.c_double,
                                      C.POINTER(C.c_double), C.c_char_p, C.c_size_t]
    rng = np.random.default_rng(3)
    rows, cols, res, xll, yll = 37, 53, 2.0, 424520.0, 565146.0
    a = rng.normal(size=(rows, cols))
    a[rng.random((rows, cols)) < 0.3] = -9999.0
    ptr = a.ctypes.data_as(C.POINTER(C.c_double))
    written = C.create_string_buffer(512)
    host.hph_raster_read.restype = C.POINTER(C.c_double)
    host.hph_raster_read.argtypes = [C.c_char_p, C.POINTER(C.c_ulong), C.POINTER(C.c_ulong), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]
    for fmt, name in (("GTiff", "a.tif"), ("ENVI", "b.bil"), ("AAIGrid", "c.asc"), ("HFA", "d.img"), ("PCIDSK", "e.pix")):
        before = host.hph_error_count()
        assert host.hph_write_raster(fmt.encode(), str(tmp_path / name).encode(), cols, rows, xll, yll, res, ptr, written, 512) == 0
        path = written.value.decode()
        if fmt == "PCIDSK":                                  # no such driver here: warned, GeoTIFF written instead
            assert path.endswith("e.tif") and host.hph_error_count() == before + 1
            assert b"not available" in host.hph_error(before)
        else:
            assert path.endswith(name) and host.hph_error_count() == before
        if path.endswith(".img"):
            # ERDAS IMAGINE: back through the reader that reads the reference's (GDAL-written) DEM
            c, r = C.c_ulong(), C.c_ulong()
            cs, x0, y0 = C.c_double(), C.c_double(), C.c_double()
            vals = host.hph_raster_read(path.encode(), C.byref(c), C.byref(r), C.byref(cs), C.byref(x0), C.byref(y0))
            assert vals and (c.value, r.value, cs.value, x0.value, y0.value) == (cols, rows, res, xll, yll)
            got = np.ctypeslib.as_array(vals, shape=(rows, cols))[::-1].copy()      # the reader returns south-first rows
            check_hfa_structure(path, cols, rows)
        elif path.endswith(".tif"):
            got, tags = read_tiff(path)
            assert tags[33550] == [res, res, 0.0] and tags[33922] == [0.0, 0.0, 0.0, xll, yll + res * rows, 0.0]
            assert tags[42113] == "-9999" and tags[34735][:4] == [1, 1, 0, 1]
        elif path.endswith(".bil"):
            got, hdr = read_envi(path)
            assert hdr["map info"] == "{Arbitrary, 1, 1, %.10g, %.10g, %.10g, %.10g}" % (xll, yll + res * rows, res, res)
            assert hdr["data ignore value"] == "-9999"
        else:
            south_first, hdr = read_asc(path)
            got = south_first[::-1]
            assert hdr["xllcorner"] == xll and hdr["yllcorner"] == yll and hdr["cellsize"] == res
        np.testing.assert_array_equal(got, a)
    assert host.hph_write_raster(b"GTiff", str(tmp_path / "no" / "dir.tif").encode(), cols, rows, xll, yll, res, ptr, written, 512) == -1


def test_output_value_derivation(host):
    # src/Datasets/CRasterDataset.cpp:185-267
    st = (C.c_double * 4)(12.5, 13.0, 1.0, -0.5)
    d = lambda name, bed=10.0: host.hph_derive_output(name.encode(), st, bed, 2.0)
    assert d("depth") == 2.5 and d("fsl") == 12.5 and d("maxdepth") == 3.0 and d("maxfsl") == 13.0
    assert d("velocityx") == 0.4 and d("velocityy") == -0.2 and d("dischargex") == 2.0
    assert abs(d("froude") - np.hypot(0.4, -0.2) / np.sqrt(9.81 * 2.5)) < 1e-15
    assert d("depth", 12.5) == -9999.0 and d("velocityx", 12.5) == -9999.0 and d("fsl", 12.5) == -9999.0


def test_configuration_is_parsed_like_the_reference(host, model_dir):
    cfg, bed_in = model_dir(scheme="MUSCL-Hancock", precision="single", duration=7200, outfreq=600)
    h = host.hph_model_load(cfg.encode(), 1)
    assert h, [host.hph_error(i) for i in range(host.hph_error_count())]
    i = info(host, h)
    assert (i["cols"], i["rows"], i["resolution"]) == (40, 30, 2.0)
    assert (i["duration"], i["output_frequency"], i["precision"], i["scheme"], i["boundaries"]) == (7200.0, 600.0, 0, 1, 3)
    courant, dry, ts = C.c_double(), C.c_double(), C.c_double()
    dyn, fric, queue = C.c_int(), C.c_int(), C.c_uint()
    host.hph_model_scheme_params(C.c_void_p(h), C.byref(courant), C.byref(dry), C.byref(ts), C.byref(dyn), C.byref(fric), C.byref(queue))
    assert (courant.value, dry.value, ts.value, dyn.value, fric.value, queue.value) == (0.5, 1e-10, 0.001, 1, 1, 16)
    st, bed, man = arrays(host, h, 30, 40)
    np.testing.assert_array_equal(bed, sc.round4(bed_in))               # 4-decimal ingestion, south-first rows
    np.testing.assert_array_equal(st[..., 0], bed)                      # depth 0 => eta = bed
    np.testing.assert_array_equal(st[..., 1], bed)                      # ... and eta_max = eta
    assert (st[..., 2:] == 0).all() and (man == 0.03).all()
    b = boundaries(host, h, 3)                                          # XML order
    assert [x["kind"] for x in b] == [0, 0, 2]
    assert b[0]["def_a"] == hc.UNIFORM_LOSS_RATE and b[0]["series"] == [0.0, 12.0, 1.0e8, 12.0] and b[0]["interval"] == 1.0e8
    assert b[1]["def_a"] == hc.UNIFORM_RAIN_INTENSITY and b[1]["interval"] == 3600.0 and b[1]["length"] == 10800.0
    assert b[2]["def_a"] == hc.DEPTH_IGNORE and b[2]["def_b"] == hc.DISCHARGE_IS_DISCHARGE and b[2]["relations"] == 3
    msgs = [host.hph_error(i).decode() for i in range(host.hph_error_count())]
    assert not any("Unrecognised" in m for m in msgs), msgs
    host.hph_model_destroy(C.c_void_p(h))


def test_bad_configurations_report_through_doError(host, model_dir, tmp_path):
    cfg, _ = model_dir()
    text = open(cfg).read()
    bad = tmp_path / "bad_scheme.xml"
    bad.write_text(text.replace('<scheme name="Godunov">', '<scheme name="lax-wendroff">'))
    assert not host.hph_model_load(str(bad).encode(), 1)
    assert "Unsupported scheme" in host.hph_error(host.hph_error_count() - 1).decode()
    bad = tmp_path / "bad_xml.xml"
    bad.write_text(text.replace("</simulation>", ""))
    assert not host.hph_model_load(str(bad).encode(), 1)
    assert "Cannot load configuration" in host.hph_error(host.hph_error_count() - 1).decode()


REF_CFG = "/root/reference/test/newcastle-centre.xml"


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present")
def test_reference_newcastle_configuration_loads_unmodified(host):
    """The reference's own test configuration (executor "OpenCL", ERDAS IMAGINE DEM with run-length
    compressed blocks) is read as it is: BASELINE.json configs[0]."""
    h = host.hph_model_load(REF_CFG.encode(), 1)
    assert h, [host.hph_error(i) for i in range(host.hph_error_count())]
    i = info(host, h)
    assert (i["cols"], i["rows"], i["resolution"], i["duration"], i["output_frequency"]) == (342, 195, 2.0, 7200.0, 600.0)
    assert (i["precision"], i["scheme"], i["boundaries"]) == (1, 0, 2)
    st, bed, man = arrays(host, h, 195, 342)
    # the .img carries its own statistics (Esta_Statistics): min 43.4375, max 81.7375, mean 56.5676
    assert bed.min() == 43.4375 and bed.max() == 81.7375 and abs(bed.mean() - 56.567615) < 1e-4
    np.testing.assert_array_equal(bed, sc.round4(bed))
    np.testing.assert_array_equal(st[..., 0], bed)
    assert (man == 0.03).all()
    b = boundaries(host, h, 2)
    assert [x["def_a"] for x in b] == [hc.UNIFORM_LOSS_RATE, hc.UNIFORM_RAIN_INTENSITY]
    assert b[1]["series"][:4] == [0.0, 70.0, 3600.0, 70.0]
    host.hph_model_destroy(C.c_void_p(h))


@pytest.mark.gpu
@pytest.mark.parametrize("scheme,name", [("Godunov", "godunov"), ("MUSCL-Hancock", "muscl-hancock"), ("Inertial", "inertial")])
def test_model_run_matches_oracle_driven_the_same_way(host, model_dir, scheme, name):
    from oracle import cpu_sim
    from tests.helpers import make_cfg
    cfg, _ = model_dir(scheme=scheme, duration=30, outfreq=10)
    h = host.hph_model_load(cfg.encode(), 0)
    assert h, [host.hph_error(i) for i in range(host.hph_error_count())]
    st0, bed, man = arrays(host, h, 30, 40)
    assert host.hph_model_run(C.c_void_p(h)) == 0
    st1, _, _ = arrays(host, h, 30, 40)
    t, dt, ok, skipped = C.c_double(), C.c_double(), C.c_uint(), C.c_uint()
    host.hph_model_clock(C.c_void_p(h), C.byref(t), C.byref(dt), C.byref(ok), C.byref(skipped))
    # the oracle, sequenced like CModel::runModel: targets at every output time, queue of 16
    ocfg = make_cfg(name, "double", 30, 40, delta=2.0, end_time=30.0)
    orc = cpu_sim.CpuSim("oracle", ocfg)
    orc.upload(st0, bed, man)
    orc.add_uniform(hc.UNIFORM_LOSS_RATE, [0.0, 1.0e8], [12.0, 12.0])
    orc.add_uniform(hc.UNIFORM_RAIN_INTENSITY, [0.0, 3600.0, 7200.0, 10800.0], [70.0, 70.0, 0.0, 0.0])
    orc.add_cell(hc.DEPTH_IGNORE, hc.DISCHARGE_IS_DISCHARGE, [10 * 40 + 1, 11 * 40 + 1, 12 * 40 + 1],
                 [[0, 0, 0, 0], [10, 0, 1.0, 0], [100000, 0, 1.0, 0]])
    for target in (10.0, 20.0, 30.0):
        orc.set_target(target)
        if orc.stats()["timestep"] <= 0.0:
            orc.update_timestep()
        while orc.stats()["time"] < target - 1e-5:
            orc.iterate(16)
    so = orc.stats()
    assert abs(t.value - so["time"]) < 1e-9 and t.value == 30.0
    assert (ok.value, skipped.value) == (so["batch_successful"], so["batch_skipped"])
    want = orc.download()
    assert np.abs(st1[..., 0] - want[..., 0]).max() <= 1e-9
    assert np.abs(st1[..., 2:] - want[..., 2:]).max() <= 1e-7
    # rasters: derived exactly as CRasterDataset::domainToRaster does, written north-first, every 10 s
    out = os.path.join(os.path.dirname(cfg), "output")
    # (derived on the device by hp_scheme_derive_raster; maxdepth is written as ERDAS IMAGINE like the reference's test config)
    from oracle import raster_oracle as ro
    out = os.path.join(os.path.dirname(cfg), "output")
    names = ["depth_%d.tif", "velX_%d.bil", "velX_%d.hdr", "fsl_%d.asc", "maxdepth_%d.img"]
    assert sorted(os.listdir(out)) == sorted(n % k for n in names for k in (10, 20, 30))
    depth, tags = read_tiff(os.path.join(out, "depth_30.tif"))
    np.testing.assert_array_equal(depth, ro.derive_raster(ro.DEPTH, st1, bed, 2.0))
    assert tags[33550][:2] == [2.0, 2.0] and tags[42113] == "-9999"
    velx, _ = read_envi(os.path.join(out, "velX_30.bil"))
    np.testing.assert_array_equal(velx, ro.derive_raster(ro.VELOCITY_X, st1, bed, 2.0))
    fsl, hdr = read_asc(os.path.join(out, "fsl_30.asc"))
    np.testing.assert_array_equal(fsl[::-1], ro.derive_raster(ro.FSL, st1, bed, 2.0))
    assert hdr["cellsize"] == 2.0 and hdr["nodata_value"] == -9999.0
    host.hph_raster_read.restype = C.POINTER(C.c_double)
    host.hph_raster_read.argtypes = [C.c_char_p, C.POINTER(C.c_ulong), C.POINTER(C.c_ulong), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]
    c, r, cs, x0, y0 = C.c_ulong(), C.c_ulong(), C.c_double(), C.c_double(), C.c_double()
    vals = host.hph_raster_read(os.path.join(out, "maxdepth_30.img").encode(), C.byref(c), C.byref(r), C.byref(cs), C.byref(x0), C.byref(y0))
    maxd = np.ctypeslib.as_array(vals, shape=(r.value, c.value))[::-1]
    np.testing.assert_array_equal(maxd, ro.derive_raster(ro.MAX_DEPTH, st1, bed, 2.0))
    check_hfa_structure(os.path.join(out, "maxdepth_30.img"), 40, 30)
    host.hph_model_destroy(C.c_void_p(h))


# ---- multi-domain sets: domains stacked north-south are merged into one ------------------------------
MULTI_XML = """<?xml version="1.0"?>
<configuration>
	<metadata><name>Two stacked domains</name><description>same schema as the reference's multi-domain sets</description></metadata>
	<execution><executor name="OpenCL"><parameter name="deviceFilter" value="GPU" /></executor></execution>
	<simulation>
		<parameter name="duration" value="{duration}" />
		<parameter name="outputFrequency" value="{duration}" />
		<parameter name="floatingPointPrecision" value="double" />
		<domainSet syncMethod="timestep">
{domains}
		</domainSet>
	</simulation>
</configuration>
"""
DOMAIN_XML = """			<domain type="cartesian" deviceNumber="{device}">
				<data sourceDir="topography/" targetDir="output/">
					<dataSource type="constant" value="depth" source="0.0" />
					<dataSource type="constant" value="manningCoefficient" source="0.030" />
					<dataSource type="raster" value="structure,dem" source="{dem}" />
					<dataTarget type="raster" value="depth" format="GTiff" target="{tag}_depth_%t.tif" />
				</data>
				<scheme name="Godunov"><parameter name="courantNumber" value="0.50" /><parameter name="queueSize" value="16" /></scheme>
				<boundaryConditions sourceDir="boundaries/">
					<timeseries type="atmospheric" name="Rainfall" value="rain-intensity" source="rainfall.csv" />
					<timeseries type="cell" name="Inflow_{tag}" depthValue="ignore" dischargeValue="volume" source="inflow.csv" mapFile="map_{tag}.csv" />
				</boundaryConditions>
			</domain>"""


def make_stacked(tmp_path, overlap_marker=0.0, duration=20):
    """A 70 x 40 terrain as two 40-row domains overlapping by 10 rows (upper listed FIRST), plus the same terrain as one
    domain.  `overlap_marker` is added to the upper file's overlap rows to show which part the merge trusts."""
    for d in ("topography", "output", "boundaries"):
        os.makedirs(tmp_path / d, exist_ok=True)
    rows, cols, res, yll = 70, 40, 2.0, 565146.0
    bed = np.round(sc.fractal_dem(rows, cols, 9, amplitude=3.0), 4)
    write_asc(tmp_path / "topography" / "full.asc", bed, res, yll=yll)
    write_asc(tmp_path / "topography" / "lower.asc", bed[:40], res, yll=yll)
    upper = bed[30:].copy()
    upper[:10] += overlap_marker
    write_asc(tmp_path / "topography" / "upper.asc", upper, res, yll=yll + 30 * res)
    (tmp_path / "boundaries" / "rainfall.csv").write_text("Time (s),Rainfall intensity (mm/hr)\n0,70\n3600,70\n7200,0\n10800,0\n")
    (tmp_path / "boundaries" / "inflow.csv").write_text("t,depth,qx,qy\n0,0,0,0\n5,0,4.0,0\n100000,0,4.0,0\n")
    (tmp_path / "boundaries" / "map_lower.csv").write_text("x,y\n1,10\n1,36\n")     # local row 36 = merged 36: the upper part's half
    (tmp_path / "boundaries" / "map_upper.csv").write_text("x,y\n2,3\n2,20\n")      # local row 3 = merged 33: the lower part's half
    (tmp_path / "boundaries" / "map_full.csv").write_text("x,y\n1,10\n")
    two = "\n".join(DOMAIN_XML.format(device=i + 1, dem=t + ".asc", tag=t) for i, t in enumerate(("upper", "lower")))
    (tmp_path / "stacked.xml").write_text(MULTI_XML.format(duration=duration, domains=two))
    (tmp_path / "single.xml").write_text(MULTI_XML.format(duration=duration, domains=DOMAIN_XML.format(device=1, dem="full.asc", tag="full")))
    return bed


def test_stacked_domains_merge_into_one(host, tmp_path):
    bed = make_stacked(tmp_path, overlap_marker=100.0)
    h = host.hph_model_load(str(tmp_path / "stacked.xml").encode(), 1)
    assert h, [host.hph_error(i) for i in range(host.hph_error_count())]
    i = info(host, h)
    assert (i["rows"], i["cols"], i["resolution"]) == (70, 40, 2.0)
    offsets = (C.c_ulong * 4)()
    host.hph_model_parts.argtypes = [C.c_void_p, C.POINTER(C.c_ulong), C.c_uint]
    assert host.hph_model_parts(C.c_void_p(h), offsets, 4) == 2 and list(offsets[:2]) == [30, 0]     # XML order: upper, lower
    st, merged, man = arrays(host, h, 70, 40)
    expect = bed.copy()
    expect[35:40] = sc.round4(bed[35:40] + 100.0)            # rows 30..34 of the overlap from the lower part, 35..39 from the upper
    np.testing.assert_array_equal(merged, expect)
    np.testing.assert_array_equal(st[..., 0], expect)
    assert (man == 0.03).all()
    # boundaries: one Rainfall for the merged domain, each cell map keeps the cells its part is authoritative for
    b = boundaries(host, h, i["boundaries"])
    assert sorted((x["kind"], x["relations"]) for x in b) == [(0, 0), (2, 1), (2, 1)]
    host.hph_model_destroy(C.c_void_p(h))
    # misaligned / non-overlapping stacks are refused like CDomainLink::canLink refuses them
    write_asc(tmp_path / "topography" / "upper.asc", bed[30:], 2.0, yll=565146.0 + 30 * 2.0 + 0.7)
    assert not host.hph_model_load(str(tmp_path / "stacked.xml").encode(), 1)
    assert b"not aligned" in host.hph_error(host.hph_error_count() - 1)
    write_asc(tmp_path / "topography" / "upper.asc", bed[30:], 2.0, yll=565146.0 + 50 * 2.0)
    assert not host.hph_model_load(str(tmp_path / "stacked.xml").encode(), 1)
    assert b"do not overlap" in host.hph_error(host.hph_error_count() - 1)


@pytest.mark.gpu
def test_stacked_domains_run_like_the_single_domain(host, tmp_path):
    """The merged run equals the run of the same terrain configured as one domain (same inflow cell), and every original
    domain gets its own raster, cropped from the merged band."""
    from oracle import raster_oracle as ro
    bed = make_stacked(tmp_path)
    (tmp_path / "boundaries" / "map_lower.csv").write_text("x,y\n1,10\n")
    (tmp_path / "boundaries" / "map_upper.csv").write_text("x,y\n")
    results = {}
    for name in ("stacked", "single"):
        h = host.hph_model_load(str(tmp_path / (name + ".xml")).encode(), 0)
        assert h, [host.hph_error(i) for i in range(host.hph_error_count())]
        assert host.hph_model_run(C.c_void_p(h)) == 0
        results[name] = arrays(host, h, 70, 40)[0]
        host.hph_model_destroy(C.c_void_p(h))
    np.testing.assert_array_equal(results["stacked"], results["single"])
    full, _ = read_tiff(str(tmp_path / "output" / "full_depth_20.tif"))
    np.testing.assert_array_equal(full, ro.derive_raster(ro.DEPTH, results["single"], bed, 2.0))
    lower, tl = read_tiff(str(tmp_path / "output" / "lower_depth_20.tif"))
    upper, tu = read_tiff(str(tmp_path / "output" / "upper_depth_20.tif"))
    np.testing.assert_array_equal(lower, full[30:])          # north-first: the lower domain is the last 40 rows
    np.testing.assert_array_equal(upper, full[:40])
    assert tl[33922][4] == 565146.0 + 40 * 2.0 and tu[33922][4] == 565146.0 + 70 * 2.0


@pytest.mark.gpu
def test_automatic_queue_and_rollback(host, model_dir):
    """queueMode="auto" with wall-clock batch sizing (CSchemeGodunov.cpp:1419-1450) changes how many iterations are
    skipped at the targets, never the result; rollbackSimulation puts the host state and clock back on the device."""
    host.hph_model_set_realtime_queue.argtypes = [C.c_void_p, C.c_int]
    host.hph_model_rollback.restype = C.c_double
    host.hph_model_rollback.argtypes = [C.c_void_p, C.c_double, C.c_double]
    host.hph_model_propose_sync.restype = C.c_double
    host.hph_model_propose_sync.argtypes = [C.c_void_p, C.c_double]
    cfg, _ = model_dir(scheme="Godunov", duration=30, outfreq=10)
    runs = []
    for realtime in (0, 1):
        h = host.hph_model_load(cfg.encode(), 0)
        assert h
        host.hph_model_set_realtime_queue(C.c_void_p(h), realtime)
        assert host.hph_model_run(C.c_void_p(h)) == 0
        t, dt, ok, skipped = C.c_double(), C.c_double(), C.c_uint(), C.c_uint()
        host.hph_model_clock(C.c_void_p(h), C.byref(t), C.byref(dt), C.byref(ok), C.byref(skipped))
        runs.append((arrays(host, h, 30, 40)[0], t.value, ok.value))
        if realtime:
            # propose a sync point from the batch statistics, then roll back to t = 12 s with a target of 50 s
            assert host.hph_model_propose_sync(C.c_void_p(h), t.value) > t.value
            new_dt = host.hph_model_rollback(C.c_void_p(h), 12.0, 50.0)
            assert new_dt > 0.0
            host.hph_model_clock(C.c_void_p(h), C.byref(t), C.byref(dt), C.byref(ok), C.byref(skipped))
            assert t.value == 12.0 and (ok.value, skipped.value) == (0, 0) and dt.value == new_dt
        host.hph_model_destroy(C.c_void_p(h))
    np.testing.assert_array_equal(runs[0][0], runs[1][0])
    assert runs[0][1] == runs[1][1] == 30.0 and runs[0][2] == runs[1][2]


def loop_state(host, h):
    syncs, outputs, method = C.c_uint(), C.c_uint(), C.c_int()
    target, last, cur = C.c_double(), C.c_double(), C.c_double()
    host.hph_model_loop_state(C.c_void_p(h), C.byref(syncs), C.byref(outputs), C.byref(target), C.byref(last), C.byref(cur), C.byref(method))
    return dict(syncs=syncs.value, outputs=outputs.value, target=target.value, last_sync=last.value, current=cur.value, sync_method=method.value)


def test_sync_method_is_parsed(host, model_dir):
    """<domainSet syncMethod=..> (src/Domain/CDomainManager.cpp:56-76); forecast is the default."""
    cfg, _ = model_dir()
    h = host.hph_model_load(cfg.encode(), 1)
    assert loop_state(host, h)["sync_method"] == 1
    host.hph_model_destroy(C.c_void_p(h))
    text = open(cfg).read().replace("<domainSet>", '<domainSet syncMethod="Timestep">')
    open(cfg, "w").write(text)
    h = host.hph_model_load(cfg.encode(), 1)
    assert loop_state(host, h)["sync_method"] == 0
    host.hph_model_destroy(C.c_void_p(h))


@pytest.mark.parametrize("duration,outfreq,syncs,targets", [
    (30, 10, [0, 10, 20, 30], [10, 20, 30, 30]),            # outputs divide the run
    (25, 10, [0, 10, 20, 25], [10, 20, 25, 25]),            # ... or not: the last target is the end of the simulation
    (30, 45, [0, 30], [30, 30]),                            # no output interval inside the run
    (7200, 600, [0, 599.999999, 600, 6600], [600, 600, 1200, 7200]),
])
def test_targets_follow_the_output_interval(host, model_dir, duration, outfreq, syncs, targets):
    """CModel::runModelUpdateTarget (src/CModel.cpp:718-770) for one domain: the end of the simulation, but never across
    the next multiple of the output frequency after the last synchronisation (:741-744)."""
    host.hph_model_next_target.restype = C.c_double
    host.hph_model_next_target.argtypes = [C.c_void_p, C.c_double]
    cfg, _ = model_dir(duration=duration, outfreq=outfreq)
    h = host.hph_model_load(cfg.encode(), 1)
    assert h
    assert [host.hph_model_next_target(C.c_void_p(h), float(t)) for t in syncs] == [float(t) for t in targets]
    host.hph_model_destroy(C.c_void_p(h))


@pytest.mark.gpu
def test_management_loop_follows_the_reference(host, model_dir):
    """CModel::runModelMain (src/CModel.cpp:1041-1139): the first pass synchronises at t = 0 and sets the first target, every
    target is the next multiple of the output frequency or the end of the simulation (:741-744), output sets are written
    only when a synchronisation lands on such a multiple (:859-866) -- a run of 25 s with outputs every 10 s syncs at 0, 10,
    20 and 25 s and writes the sets of 10 and 20 s -- and timestep synchronisation falls back to forecast for one domain
    (:503-505)."""
    from oracle import cpu_sim
    from tests.helpers import make_cfg
    cfg, _ = model_dir(scheme="Godunov", duration=25, outfreq=10)
    text = open(cfg).read().replace("<domainSet>", '<domainSet syncMethod="timestep">')
    open(cfg, "w").write(text)
    h = host.hph_model_load(cfg.encode(), 0)
    assert h, [host.hph_error(i) for i in range(host.hph_error_count())]
    st0, bed, man = arrays(host, h, 30, 40)
    assert host.hph_model_run(C.c_void_p(h)) == 0
    ls = loop_state(host, h)
    assert ls == dict(syncs=4, outputs=2, target=25.0, last_sync=25.0, current=25.0, sync_method=1)
    out = os.path.join(os.path.dirname(cfg), "output")
    assert sorted(f for f in os.listdir(out) if f.startswith("depth_")) == ["depth_10.tif", "depth_20.tif"]
    # the same sequence of targets on the oracle: identical iteration counts and state
    t, dt, ok, skipped = C.c_double(), C.c_double(), C.c_uint(), C.c_uint()
    host.hph_model_clock(C.c_void_p(h), C.byref(t), C.byref(dt), C.byref(ok), C.byref(skipped))
    st1, _, _ = arrays(host, h, 30, 40)
    ocfg = make_cfg("godunov", "double", 30, 40, delta=2.0, end_time=25.0)
    orc = cpu_sim.CpuSim("oracle", ocfg)
    orc.upload(st0, bed, man)
    orc.add_uniform(hc.UNIFORM_LOSS_RATE, [0.0, 1.0e8], [12.0, 12.0])
    orc.add_uniform(hc.UNIFORM_RAIN_INTENSITY, [0.0, 3600.0, 7200.0, 10800.0], [70.0, 70.0, 0.0, 0.0])
    orc.add_cell(hc.DEPTH_IGNORE, hc.DISCHARGE_IS_DISCHARGE, [10 * 40 + 1, 11 * 40 + 1, 12 * 40 + 1],
                 [[0, 0, 0, 0], [10, 0, 1.0, 0], [100000, 0, 1.0, 0]])
    for target in (10.0, 20.0, 25.0):
        orc.set_target(target)
        if orc.stats()["timestep"] <= 0.0:
            orc.update_timestep()
        while orc.stats()["time"] < target - 1e-5:
            orc.iterate(16)
    so = orc.stats()
    assert t.value == 25.0 and (ok.value, skipped.value) == (so["batch_successful"], so["batch_skipped"])
    assert np.abs(st1[..., 0] - orc.download()[..., 0]).max() <= 1e-9
    orc.close()
    host.hph_model_destroy(C.c_void_p(h))


# ---- decomposed models from the configuration: <domain deviceNumber=..> stacks -> the strip engine --------
def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("scheme,exchange", [("Godunov", "nccl"), ("MUSCL-Hancock", "nccl"), ("Inertial", "nccl"), ("MUSCL-Hancock", "peer"),
                                             ("Godunov", "peer")])
def test_stacked_domains_run_as_row_strips_on_their_devices(host, tmp_path, scheme, exchange, monkeypatch):
    """A two-<domain> configuration naming deviceNumber 1 and 2 runs as two row strips, one per GPU (NCCL halo exchange and
    dt all-reduce inside the library -- or, with HIPIMS_STRIP_EXCHANGE=peer, the peer-memory exchange kernel and graph
    replay -- one host thread per strip), bit-identical to the same terrain configured as one
    domain on one GPU -- states, clock and per-domain rasters (src/Domain/CDomainManager.cpp:56-282,
    Links/CDomainLink.cpp:286-382 re-targeted)."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.setenv("HIPIMS_STRIP_EXCHANGE", exchange)
    from oracle import raster_oracle as ro
    bed = make_stacked(tmp_path, duration=20)
    (tmp_path / "boundaries" / "map_lower.csv").write_text("x,y\n1,10\n")
    (tmp_path / "boundaries" / "map_upper.csv").write_text("x,y\n")
    for name in ("stacked.xml", "single.xml"):
        p = tmp_path / name
        p.write_text(p.read_text().replace('<scheme name="Godunov">', '<scheme name="%s">' % scheme))
    host.hph_model_strips.argtypes = [C.c_void_p]
    host.hph_model_strips.restype = C.c_uint
    results, clocks = {}, {}
    for name in ("stacked", "single"):
        h = host.hph_model_load(str(tmp_path / (name + ".xml")).encode(), 0)
        assert h, [host.hph_error(i) for i in range(host.hph_error_count())]
        assert host.hph_model_strips(C.c_void_p(h)) == (2 if name == "stacked" else 1)
        assert host.hph_model_run(C.c_void_p(h)) == 0, [host.hph_error(i) for i in range(host.hph_error_count())]
        results[name] = arrays(host, h, 70, 40)[0]
        t, dt, ok, skipped = C.c_double(), C.c_double(), C.c_uint(), C.c_uint()
        host.hph_model_clock(C.c_void_p(h), C.byref(t), C.byref(dt), C.byref(ok), C.byref(skipped))
        clocks[name] = (t.value, dt.value, ok.value, skipped.value)
        host.hph_model_destroy(C.c_void_p(h))
    assert not np.array_equal(results["single"][..., 0], bed)                 # it rained and the inflow ran
    np.testing.assert_array_equal(results["stacked"], results["single"])
    assert clocks["stacked"] == clocks["single"]
    full, _ = read_tiff(str(tmp_path / "output" / "full_depth_20.tif"))
    np.testing.assert_array_equal(full, ro.derive_raster(ro.DEPTH, results["single"], bed, 2.0))
    lower, _ = read_tiff(str(tmp_path / "output" / "lower_depth_20.tif"))
    upper, _ = read_tiff(str(tmp_path / "output" / "upper_depth_20.tif"))
    np.testing.assert_array_equal(lower, full[30:])
    np.testing.assert_array_equal(upper, full[:40])


@pytest.mark.gpu
def test_single_domain_spread_over_devices_on_request(host, model_dir):
    """--devices / hph_model_load_on: one <domain> split into row strips over every GPU of the box."""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    cfg, bed = model_dir(scheme="MUSCL-Hancock", duration=30, outfreq=30)
    host.hph_model_load_on.restype = C.c_void_p
    host.hph_model_load_on.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_int]
    host.hph_model_strips.argtypes = [C.c_void_p]
    host.hph_model_strips.restype = C.c_uint
    n = min(n, 4)
    out = {}
    for devices in ([1], list(range(1, n + 1))):
        arr = (C.c_int * len(devices))(*devices)
        h = host.hph_model_load_on(cfg.encode(), arr, len(devices))
        assert h, [host.hph_error(i) for i in range(host.hph_error_count())]
        assert host.hph_model_strips(C.c_void_p(h)) == len(devices)
        assert host.hph_model_run(C.c_void_p(h)) == 0, [host.hph_error(i) for i in range(host.hph_error_count())]
        out[len(devices)] = arrays(host, h, 30, 40)[0]
        host.hph_model_destroy(C.c_void_p(h))
    np.testing.assert_array_equal(out[n], out[1])
