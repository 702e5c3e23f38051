"""createdatamatrix (row scatter) and the result writers: oracle / host logic on CPU, device scatter on GPU."""
import numpy as np
import pytest

from carparkingmaps_b200 import results_io
from carparkingmaps_b200.synth import synth_city
from oracle import oracle as O

T = 24


def _rows_from_dense(dm, rs, shuffle=True, dup=0):
    """CSV-style columns of the non-empty cells: zone N is written as ID 0 and hour 24 as 0, as in the Uber files
    (src/createdatamatrix.jl:9-17 maps them back)."""
    N = dm.shape[0]
    i, j, t = np.nonzero(dm[:, :, :, 0])
    order = rs.permutation(len(i)) if shuffle else np.arange(len(i))
    i, j, t = i[order], j[order], t[order]
    src = (i + 1) % N
    dst = (j + 1) % N
    hod = (t + 1) % 24
    mean, sd = dm[i, j, t, 0].copy(), dm[i, j, t, 1].copy()
    if dup:   # duplicate keys with other values EARLIER in the file: the last row must win
        k = rs.choice(len(i), dup, replace=False)
        src = np.concatenate([src[k], src]); dst = np.concatenate([dst[k], dst]); hod = np.concatenate([hod[k], hod])
        mean = np.concatenate([mean[k] + 1000.0, mean]); sd = np.concatenate([sd[k] + 7.0, sd])
    return src, dst, hod, mean, sd


def test_oracle_createdatamatrix_semantics():
    rs = np.random.default_rng(0)
    N = 9
    dm, _ = synth_city(N, T, seed=2, density=0.3)
    src, dst, hod, mean, sd = _rows_from_dense(dm, rs, dup=20)
    got, bad = O.createdatamatrix(src, dst, hod, mean, sd, N, T=T)
    assert bad == 0
    np.testing.assert_array_equal(got, dm)                     # ID 0 -> N, hour 0 -> 24, last row wins
    # literal loop restatement (src/createdatamatrix.jl:8-23)
    ref = np.zeros_like(dm)
    for r in range(len(src)):
        a, b, h = int(src[r]) or N, int(dst[r]) or N, int(hod[r]) or 24
        ref[a - 1, b - 1, h - 1, 0] = mean[r]; ref[a - 1, b - 1, h - 1, 1] = sd[r]
    np.testing.assert_array_equal(got, ref)
    _, bad = O.createdatamatrix([N + 1, 1], [1, 1], [1, 25], [5.0, 6.0], [1.0, 1.0], N, T=T)
    assert bad == 2                                            # out of range: BoundsError in the reference, skipped here


@pytest.mark.gpu
@pytest.mark.parametrize("N,density,dup", [(9, 0.3, 20), (64, 0.09, 500), (40, 1.0, 0)])
def test_device_row_scatter_matches_oracle(ctx, N, density, dup):
    from carparkingmaps_b200 import City
    rs = np.random.default_rng(N)
    dm, dist = synth_city(N, T, seed=N, density=density)
    src, dst, hod, mean, sd = _rows_from_dense(dm, rs, dup=dup)
    src = np.concatenate([src, [N + 5]]); dst = np.concatenate([dst, [1]]); hod = np.concatenate([hod, [3]])
    mean = np.concatenate([mean, [1.0]]); sd = np.concatenate([sd, [1.0]])
    city = City(ctx, N, T)
    skipped = city.upload_rows(src, dst, hod, mean, sd, dist)
    want, bad = O.createdatamatrix(src, dst, hod, mean, sd, N, T=T)
    assert skipped == bad == 1
    dm_d, dist_d = city.download_inputs()
    np.testing.assert_array_equal(dm_d, want)
    np.testing.assert_array_equal(dm_d, dm)
    np.testing.assert_array_equal(dist_d, dist)
    city.close()


def test_julia_float_formatting():
    cases = {1.0: "1.0", 0.5: "0.5", 1e-5: "1.0e-5", 1.5e-7: "1.5e-7", 0.0001: "0.0001", 123456.789: "123456.789",
             1e21: "1.0e21", 2.5e22: "2.5e22", 0.0: "0.0", 1e6: "1.0e6", 2357000.0: "2.357e6", 1e15: "1.0e15", 999999.5: "999999.5",
             100000.0: "100000.0", -2400000.0: "-2.4e6", 1234567.891: "1.234567891e6", -3.25e-9: "-3.25e-9", float("nan"): "NaN", 0.00012: "0.00012",
             0.00004166666666666667: "4.1666666666666665e-5" if repr(0.00004166666666666667) == "4.1666666666666665e-05" else None}
    for x, want in cases.items():
        if want is not None:
            assert results_io.jl_float(x) == want, (x, results_io.jl_float(x), want)
    for x in np.random.default_rng(1).random(200) * 10.0 ** np.random.default_rng(2).integers(-8, 12, 200):
        assert float(results_io.jl_float(x)) == x             # round-trips


def test_result_files_have_the_reference_format(tmp_path):
    N = 5
    rs = np.random.default_rng(3)
    parking = rs.random((N, T)) / N
    traffic = rs.random(T)
    path = results_io.createresultsdirectory(str(tmp_path), "Melbourne")
    results_io.createresultsdirectory(str(tmp_path), "Melbourne")          # tolerates an existing folder (:8-13)
    results_io.saveresults(path, "melbourne-2017-1-OnlyWeekdays.csv", parking, traffic)
    results_io.saveparameters(path, 24, N, 1000, N * 1000, 0.5, 0.1, 0.9, 2, 0.0712)
    lines = open(tmp_path / "Melbourne" / "results_parkingdensities_melbourne-2017-1-OnlyWeekdays.csv").read().splitlines()
    assert lines[0].split(",") == [f"t = {t}h" for t in range(1, 25)] and len(lines) == N + 1    # src/saveresults.jl:35-38
    np.testing.assert_array_equal(np.array([[float(v) for v in l.split(",")] for l in lines[1:]]), parking)
    lines = open(tmp_path / "Melbourne" / "results_trafficactivity_melbourne-2017-1-OnlyWeekdays.csv").read().splitlines()
    assert len(lines) == 2 and [float(v) for v in lines[1].split(",")] == list(traffic)
    lines = open(tmp_path / "Melbourne" / "sampling_parameters.csv").read().splitlines()
    assert lines[0].split(",") == results_io.PARAMETER_HEADER and lines[1].startswith("24.0,5.0,1000.0,5000.0,0.5,0.1,0.9,2.0,0.0712")


def test_oracle_distance_matrix_formula():
    # src/processgeodata.jl:150-166 with Melbourne-like centroids
    rs = np.random.default_rng(5)
    lat, lon = -37.8 + 0.3 * rs.random(12), 144.9 + 0.4 * rs.random(12)
    d = O.distancematrix(lat, lon)
    assert (np.diag(d) == 1).all() and np.array_equal(d, d.T)
    i, j = 2, 7
    want = 111.3 * np.sqrt(np.cos((lat[i] + lat[j]) / 2 * 0.01745) ** 2 * (lon[i] - lon[j]) ** 2 + (lat[i] - lat[j]) ** 2)
    np.testing.assert_allclose(d[i, j], want, rtol=1e-14)
    assert 0 < d[i, j] < 60          # a city-sized box


@pytest.mark.gpu
def test_device_distance_matrix_matches_oracle(ctx):
    from carparkingmaps_b200 import api
    rs = np.random.default_rng(6)
    for N in (1, 7, 300):
        lat, lon = -37.8 + 0.3 * rs.random(N), 144.9 + 0.4 * rs.random(N)
        got = api.distancematrix(lat, lon)
        want = O.distancematrix(lat, lon)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)     # device cos() vs libm: a few ulp
        assert (np.diag(got) == 1).all() and np.array_equal(got, got.T)


# ---------------------------------------------------------------------------------------------------------
# polygon area / centroid block of processgeodata (src/processgeodata.jl:99-148)
# ---------------------------------------------------------------------------------------------------------
def _polygon_matrices(rs, N, V, *, duplicate=True):
    """Zero-terminated vertex lists the way the reference's GeoJSON walk fills them: a ring around a Melbourne-like
    centre, every vertex stored twice (src/processgeodata.jl:44-47 runs once per coordinate of the pair)."""
    lon, lat = np.zeros((N, V), order="F"), np.zeros((N, V), order="F")
    rings = []
    for i in range(N):
        if i % 7 == 3:
            rings.append(None)              # a zone id that never occurs in the file: row stays zero
            continue
        k = int(rs.integers(3, max(4, min(40, V // 2 - 2))))   # the doubled, closed ring and its terminator fit in V
        ang = np.sort(rs.random(k)) * 2 * np.pi
        rad = 0.01 + 0.02 * rs.random(k)
        x = 144.9 + 0.4 * rs.random() + rad * np.cos(ang)
        y = -37.8 + 0.3 * rs.random() + rad * np.sin(ang)
        x, y = np.append(x, x[0]), np.append(y, y[0])     # GeoJSON rings repeat the first vertex at the end
        rings.append((x.copy(), y.copy()))
        if duplicate:
            x, y = np.repeat(x, 2), np.repeat(y, 2)
        lon[i, :len(x)], lat[i, :len(y)] = x, y
    return lon, lat, rings


def test_oracle_polygon_centroids_match_the_shoelace_formulas():
    rs = np.random.default_rng(11)
    lon, lat, rings = _polygon_matrices(rs, 20, 200)
    lon0, lat0 = lon.copy(), lat.copy()
    c_lat, c_lon, area = O.polygoncentroids(lon, lat)
    assert np.array_equal(lon, lon0) and np.array_equal(lat, lat0)        # the wrapper works on copies
    for i, ring in enumerate(rings):
        if ring is None:
            assert c_lat[i] == 0 and c_lon[i] == 0 and area[i] == 0       # :105 skips zones without vertices
            continue
        x, y = ring
        cross = x[:-1] * y[1:] - x[1:] * y[:-1]
        A = cross.sum() / 2                                               # standard signed area (x = lon, y = lat)
        cx = ((x[:-1] + x[1:]) * cross).sum() / (6 * A)
        cy = ((y[:-1] + y[1:]) * cross).sum() / (6 * A)
        np.testing.assert_allclose(area[i], -A, rtol=1e-6)                # the reference's sum is lat*lon' - lat'*lon
        # the sums are differences of products of ~1e4-sized numbers: ~1e-7 degrees is what fp64 leaves
        np.testing.assert_allclose(c_lon[i], cx, atol=1e-5)
        np.testing.assert_allclose(c_lat[i], cy, atol=1e-5)
        assert x.min() < c_lon[i] < x.max() and y.min() < c_lat[i] < y.max()


def test_oracle_polygon_duplicate_vertices_do_not_move_the_centroid():
    rs = np.random.default_rng(12)
    lon2, lat2, _ = _polygon_matrices(rs, 10, 200, duplicate=True)
    rs = np.random.default_rng(12)
    lon1, lat1, _ = _polygon_matrices(rs, 10, 200, duplicate=False)
    a, b = O.polygoncentroids(lon2, lat2), O.polygoncentroids(lon1, lat1)
    for u, v in zip(a, b):
        np.testing.assert_allclose(u, v, rtol=0, atol=1e-9)               # a repeated vertex adds exact zeros


@pytest.mark.gpu
def test_device_polygon_centroids_match_oracle(ctx):
    from carparkingmaps_b200 import api
    rs = np.random.default_rng(13)
    for N, V in ((1, 50), (9, 120), (300, 90)):
        lon, lat, _ = _polygon_matrices(rs, N, V)
        if N == 9:
            lon[5, 4] = 0.0                   # an embedded zero longitude: the area pass stops there, the centroid pass does not
        lon0 = lon.copy()
        got = api.polygoncentroids(lon, lat)
        want = O.polygoncentroids(lon, lat)
        assert np.array_equal(lon, lon0)      # host inputs untouched
        for g, w in zip(got, want):
            assert np.array_equal(g, w, equal_nan=True)                   # same operations in the same order: bit-exact
    # a row without a terminator inside the window: both sides stop at the last column
    lon, lat, _ = _polygon_matrices(rs, 4, 30)
    lon[0, :], lat[0, :] = 145.0 + 0.01 * rs.random(30), -37.8 + 0.01 * rs.random(30)
    got, want = api.polygoncentroids(lon, lat), O.polygoncentroids(lon, lat)
    for g, w in zip(got, want):
        assert np.array_equal(g, w, equal_nan=True)
