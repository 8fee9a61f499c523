"""PostgreSQL surface: `COPY`-loadable CSV for the reference's tables in the exact column
order of db/schema.sql, plus the DDL of the (new) `psms` table.  The in-HBM index replaces the database on the hot
path; these exporters restore it as the system of record for downstream tools.  No PostgreSQL server exists in the
build image, so the tests check the CSV bytes.

  proteins           id, accession, header, aa_sequence, is_completely_digested                 (db/schema.sql:2-8)
  peptides / decoys  id, aa_sequence, length, number_of_missed_cleavages, weight, r_count ... y_count, a_count
                                                                                          (db/schema.sql:14-43,163-192)
  peptides_proteins  peptide_id, protein_id                                                     (db/schema.sql:156-160)
"""
import csv
import io
import re

from ._abi import ALPHABET

# schema order of the count columns: r n d c e q g h j k m f p o s t u v w y, and a LAST (db/schema.sql:20-40)
COUNT_COLUMNS = "rndceqghjkmfpostuvwya"
_COUNT_INDEX = [ALPHABET.index(c.upper()) for c in COUNT_COLUMNS]
# models/protein.rs:27
ACCESSION_RE = re.compile(r"[OPQ][0-9][A-Z0-9]{3}[0-9]|[A-NR-Z][0-9]([A-Z][A-Z0-9]{2}[0-9]){1,2}")

PSMS_DDL = """-- psms: new table (the reference has none: scoring happened in an external Comet run)
CREATE TABLE psms (
    spectrum_id TEXT NOT NULL,
    scan_id TEXT NOT NULL,
    rank SMALLINT NOT NULL,
    is_decoy BOOLEAN NOT NULL,
    peptide_id BIGINT,
    aa_sequence VARCHAR(60) NOT NULL,
    modres TEXT NOT NULL,
    weight BIGINT NOT NULL,
    precursor_mass BIGINT NOT NULL,
    charge SMALLINT NOT NULL,
    score REAL NOT NULL,
    n_candidates INTEGER NOT NULL,
    PRIMARY KEY (spectrum_id, rank)
);
"""


def extract_accession(header):
    """Protein::extract_accession_from_header (models/protein.rs:24-35): first match of the UniProt pattern."""
    m = ACCESSION_RE.search(header)
    return m.group(0) if m else ""


def _csv(rows):
    buf = io.StringIO()
    w = csv.writer(buf, lineterminator="\n")
    w.writerows(rows)
    return buf.getvalue()


def proteins_csv(headers, sequences):
    return _csv([(i + 1, extract_accession(h), h, s, "t") for i, (h, s) in enumerate(zip(headers, sequences))])


def peptides_csv(table, sequences=None, id_base=1):
    """`table` = Engine.peptides(); peptide id = row index + id_base (the ids the PSM rows carry)."""
    from .api import Engine
    seqs = sequences if sequences is not None else Engine.sequences_of(table)
    rows = []
    for k, s in enumerate(seqs):
        c = table["counts"][k]
        rows.append([k + id_base, s, len(s), int(table["missed_cleavages"][k]), int(table["weight"][k])] + [int(c[i]) for i in _COUNT_INDEX])
    return _csv(rows)


def peptides_proteins_csv(table, id_base=1):
    rows = []
    ao, ap = table["assoc_off"], table["assoc_protein"]
    for k in range(len(ao) - 1):
        for j in range(int(ao[k]), int(ao[k + 1])):
            rows.append((k + id_base, int(ap[j]) + 1))
    return _csv(rows)


def decoys_csv(decoy_table, id_base=1):
    """`decoy_table` = Engine.generate_decoys(...) / last_decoys(); rows of table `decoys` (unique by sequence:
    UNIQUE (aa_sequence, weight), db/schema.sql:190).  Decoy::new: weight = unmodified, missed cleavages 0."""
    raw, so = decoy_table["seq"].tobytes(), decoy_table["seq_off"]
    seen, rows = set(), []
    for i in range(len(so) - 1):
        s = raw[int(so[i]):int(so[i + 1])].decode()
        if s in seen:
            continue
        seen.add(s)
        rows.append([len(rows) + id_base, s, len(s), 0, int(decoy_table["weight"][i])] + [s.count(c.upper()) for c in COUNT_COLUMNS])
    return _csv(rows)


def psms_csv(psms, ids, sequences_of_candidate, modres_of_candidate, precursor_masses):
    """Rows of `psms` from the PSM table of Engine.identify.  ids[s] = (spectrum_id, scan_id);
    sequences_of_candidate(s, row) / modres_of_candidate(s, row) resolve a PSM row to its sequence / ModRes string."""
    rows = []
    n, k = psms.shape
    for s in range(n):
        for r in range(k):
            row = psms[s, r]
            if int(row["rank"]) == 0:
                continue
            rows.append((ids[s][0], ids[s][1], int(row["rank"]), "t" if row["is_decoy"] else "f",
                         "" if row["is_decoy"] else int(row["candidate"]), sequences_of_candidate(s, row), modres_of_candidate(s, row),
                         int(row["mod_weight"]), int(precursor_masses[s]), int(row["charge"]), repr(float(row["score"])),
                         int(row["n_targets"]) + int(row["n_decoys"])))
    return _csv(rows)
