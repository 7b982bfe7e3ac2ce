"""Synthetic KDD99-shaped and CICIDS2017-shaped raw flow records.

The real CSVs are not available (no network); generators are seeded torch code that runs on
whatever device it is given, so the bench generates on the GPU and tests can copy to the host.
Records are AoS: KDD = 42 x 4 B = 168 B (38 f32 numerics, 3 int32 dictionary codes, int32 label
code, in the file order of kdd99.py:15-23); CICIDS = 78 f32 + int32 label code = 316 B.
"""
import numpy as np
import torch

from .encode import RecordSchema

KDD_COLUMNS = ["duration", "protocol_type", "service", "flag", "src_bytes", "dst_bytes", "land", "wrong_fragment",
               "urgent", "hot", "num_failed_logins", "logged_in", "num_compromised", "root_shell", "su_attempted",
               "num_root", "num_file_creations", "num_shells", "num_access_files", "num_outbound_cmds", "is_host_login",
               "is_guest_login", "count", "srv_count", "serror_rate", "srv_serror_rate", "rerror_rate", "srv_rerror_rate",
               "same_srv_rate", "diff_srv_rate", "srv_diff_host_rate", "dst_host_count", "dst_host_srv_count",
               "dst_host_same_srv_rate", "dst_host_diff_srv_rate", "dst_host_same_src_port_rate",
               "dst_host_srv_diff_host_rate", "dst_host_serror_rate", "dst_host_srv_serror_rate", "dst_host_rerror_rate",
               "dst_host_srv_rerror_rate", "label"]
KDD_CATEGORICAL = ["protocol_type", "service", "flag"]
KDD_RATE = set(KDD_COLUMNS[24:31] + KDD_COLUMNS[33:41])
KDD_BINARY = {"land", "logged_in", "root_shell", "is_host_login", "is_guest_login"}

# KDD99-full label counts (Appendix B) -> class priors of the 23 latent attack types
KDD_LABELS = ["smurf", "neptune", "normal", "satan", "ipsweep", "portsweep", "nmap", "back", "warezclient", "teardrop",
              "pod", "guess_passwd", "buffer_overflow", "land", "warezmaster", "imap", "rootkit", "loadmodule",
              "ftp_write", "multihop", "phf", "perl", "spy"]
KDD_COUNTS = [2807886, 1072017, 972781, 15892, 12481, 10413, 2316, 2203, 1020, 979, 264, 53, 30, 21, 20, 12, 10, 9, 8, 7,
              4, 3, 2]
_FIVE = {"DoS": ["back", "land", "neptune", "pod", "smurf", "teardrop"], "Probe": ["ipsweep", "nmap", "portsweep", "satan"],
         "R2L": ["ftp_write", "guess_passwd", "imap", "multihop", "phf", "spy", "warezclient", "warezmaster"],
         "U2R": ["buffer_overflow", "loadmodule", "perl", "rootkit"], "normal": ["normal"]}


def kdd_schema():
    return RecordSchema([(c, "code" if c in KDD_CATEGORICAL or c == "label" else "f32") for c in KDD_COLUMNS])


def _class_map(n_classes):
    if n_classes == 23:
        return list(KDD_LABELS), list(range(23))
    if n_classes == 5:
        names = sorted(_FIVE)
        return names, [next(i for i, k in enumerate(names) if l in _FIVE[k]) for l in KDD_LABELS]
    if n_classes == 2:
        return ["attack", "normal"], [1 if l == "normal" else 0 for l in KDD_LABELS]
    raise ValueError("n_classes must be 2, 5 or 23")


def make_kdd(n, n_classes=5, seed=2019, device="cpu", label_noise=0.01, row_offset=0):
    """-> (records uint8 [n,168], dictionaries {column: [strings]}) ; dictionary codes are in a scrambled
    order (not frequency order) so the StringIndexer LUT is exercised."""
    dev = torch.device(device)
    g = torch.Generator(device=dev); g.manual_seed(int(seed) + 7919 * int(row_offset))
    pg = np.random.default_rng(12345)                     # class-conditional parameters: fixed, host side
    L = 23
    pri = torch.tensor(np.asarray(KDD_COUNTS, np.float64) / sum(KDD_COUNTS), device=dev)
    z = torch.multinomial(pri, n, replacement=True, generator=g)            # latent attack type
    names, cmap = _class_map(n_classes)
    cmap_t = torch.tensor(cmap, device=dev)
    y = cmap_t[z]
    flip = torch.rand(n, device=dev, generator=g) < label_noise
    y = torch.where(flip, torch.randint(0, len(names), (n,), device=dev, generator=g), y)

    def cat_column(K, conc):
        # per latent type: a peaked distribution over K categories
        probs = np.zeros((L, K))
        for l in range(L):
            w = 1.0 / np.arange(1, K + 1) ** conc
            probs[l] = pg.permutation(w) if l > 2 else np.roll(w, l * 2)
            probs[l] /= probs[l].sum()
        p = torch.tensor(probs, device=dev)[z]                               # [n, K]
        u = torch.rand(n, 1, device=dev, generator=g, dtype=torch.float64)
        return (p.cumsum(1) < u).sum(1).clamp_(max=K - 1).to(torch.int32)

    cols = {}
    dicts = {}
    for name, K, conc in (("protocol_type", 3, 3.0), ("service", 70, 2.2), ("flag", 11, 2.5)):
        rank_like = cat_column(K, conc)
        scramble = pg.permutation(K)                                          # code = scramble[category]
        cols[name] = torch.tensor(scramble, device=dev, dtype=torch.int32)[rank_like.long()]
        inv = np.argsort(scramble)
        base = {"protocol_type": ["icmp", "tcp", "udp"]}.get(name, ["%s_%02d" % (name[:3], i) for i in range(K)])
        dicts[name] = [base[inv[c]] for c in range(K)]
    lab_scramble = pg.permutation(len(names))
    cols["label"] = torch.tensor(lab_scramble, device=dev, dtype=torch.int32)[y]
    dicts["label"] = [names[np.argsort(lab_scramble)[c]] for c in range(len(names))]

    loc = torch.tensor(pg.uniform(0.0, 1.0, (L, len(KDD_COLUMNS))), device=dev, dtype=torch.float32)
    smurf_like = z < 2                                                        # smurf / neptune: near-identical rows
    for j, name in enumerate(KDD_COLUMNS):
        if name in cols:
            continue
        u = torch.rand(n, device=dev, generator=g)
        m = loc[:, j][z]
        if name in KDD_RATE:
            v = torch.where(smurf_like, (m > 0.5).float(), (0.6 * m + 0.4 * u).clamp(0, 1))
            v = torch.round(v * 100.0) / 100.0
        elif name in KDD_BINARY:
            v = (u < 0.05 + 0.9 * (m > 0.7).float()).float()
        elif name == "num_outbound_cmds":
            v = torch.zeros(n, device=dev)
        elif name in ("count", "srv_count"):
            v = torch.where(smurf_like, torch.full_like(u, 511.0), torch.floor(m * 300.0 * u + 1.0))
        elif name in ("dst_host_count", "dst_host_srv_count"):
            v = torch.where(smurf_like, torch.full_like(u, 255.0), torch.floor(255.0 * (0.5 * m + 0.5 * u)))
        elif name in ("src_bytes", "dst_bytes", "duration"):
            heavy = torch.floor(torch.exp(u * (4.0 + 14.0 * m)))              # log-uniform heavy tail
            v = torch.where(smurf_like, torch.floor(1000.0 * m + 32.0), heavy - 1.0)
        else:
            v = torch.floor(torch.clamp(-torch.log(1.0 - u * 0.999) * 2.0 * m - 1.5, min=0.0))   # mostly zero counts
        cols[name] = v.to(torch.float32)
    schema = kdd_schema()
    rec = torch.empty((n, schema.row_bytes), dtype=torch.uint8, device=dev)
    rec32 = rec.view(torch.int32)
    for j, name in enumerate(KDD_COLUMNS):
        c = cols[name]
        rec32[:, j] = c if c.dtype == torch.int32 else c.view(torch.int32)
    return rec, dicts


CICIDS_LABELS = ["BENIGN", "DoS Hulk", "PortScan", "DDoS", "DoS GoldenEye", "FTP-Patator", "SSH-Patator", "DoS slowloris",
                 "DoS Slowhttptest", "Bot", "Web Attack Brute Force", "Web Attack XSS", "Infiltration",
                 "Web Attack Sql Injection", "Heartbleed"]
CICIDS_COUNTS = [2273097, 231073, 158930, 128027, 10293, 7938, 5897, 5796, 5499, 1966, 1507, 652, 36, 21, 11]


def cicids_schema(n_features=78, dtype="f32"):
    """dtype="f64": the layout Spark's inferSchema gives the CICIDS2017 CSVs (cicids17.py:19-20 -> DoubleType columns):
    78 x 8 + 4 = 628 B per record."""
    return RecordSchema([("f%02d" % i, dtype) for i in range(n_features)] + [("Label", "code")])


def make_cicids(n, n_classes=15, seed=2019, device="cpu", label_noise=0.01, nan_fraction=0.0, n_features=78, dtype="f32",
                row_offset=0):
    """-> (records uint8 [n, row_bytes], dictionaries {"Label": [...]}).  Mix of integer counters, µs durations,
    many-digit rates, 8 constant-zero columns; optional NaN in two rate columns (handleInvalid='skip').  dtype="f64" keeps
    the rate columns' full double precision (more than 7 significant digits: an f32 copy of the same record can land on the
    other side of a split threshold)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev); g.manual_seed(int(seed) + 7919 * int(row_offset))
    pg = np.random.default_rng(54321)
    L = n_classes
    real = torch.float64 if dtype == "f64" else torch.float32
    pri = np.asarray(CICIDS_COUNTS[:L], np.float64); pri /= pri.sum()
    z = torch.multinomial(torch.tensor(pri, device=dev), n, replacement=True, generator=g)
    flip = torch.rand(n, device=dev, generator=g) < label_noise
    y = torch.where(flip, torch.randint(0, L, (n,), device=dev, generator=g), z)
    loc = torch.tensor(pg.uniform(0.0, 1.0, (L, n_features)), device=dev, dtype=real)
    informative = set(pg.choice(n_features, min(14, n_features // 2), replace=False).tolist())
    zero_cols = set(pg.choice([i for i in range(n_features) if i not in informative], min(8, n_features // 5), replace=False).tolist())
    schema = cicids_schema(n_features, dtype)
    rec = torch.empty((n, schema.row_bytes), dtype=torch.uint8, device=dev)
    rec32 = rec.view(torch.int32)
    wpf = 2 if dtype == "f64" else 1                                          # 32-bit words per feature field
    for j in range(n_features):
        u = torch.rand(n, device=dev, generator=g, dtype=real)
        m = loc[:, j][z] if j in informative else torch.full((n,), 0.5, device=dev, dtype=real)
        if j in zero_cols:
            v = torch.zeros(n, device=dev, dtype=real)
        elif j % 3 == 0:
            v = torch.floor(torch.exp(u * (3.0 + 15.0 * m)))                  # counters / µs durations up to ~1e8
        elif j % 3 == 1:
            v = (u * m * 1.0e6 + u * u * 37.0) / (1.0 + 3.0 * m)              # rates with many significant digits
        else:
            v = torch.floor(u * (40.0 * m + 2.0))
        if nan_fraction > 0 and j in (14, 15):
            v = torch.where(torch.rand(n, device=dev, generator=g) < nan_fraction, torch.full_like(v, float("nan")), v)
        if dtype == "f64":
            rec32[:, 2 * j:2 * j + 2] = v.to(torch.float64).view(torch.int32).reshape(n, 2)
        else:
            rec32[:, j] = v.to(torch.float32).view(torch.int32)
    scramble = pg.permutation(L)
    rec32[:, wpf * n_features] = torch.tensor(scramble, device=dev, dtype=torch.int32)[y]
    names = CICIDS_LABELS[:L]
    return rec, {"Label": [names[np.argsort(scramble)[c]] for c in range(L)]}
