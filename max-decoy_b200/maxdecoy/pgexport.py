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


# ------------------------------------------------------------------------------------------------
# .env / PGSQL_URL (utility/database_connection.rs:8-20) and the way back from the database
# ------------------------------------------------------------------------------------------------
def database_url(env_path=".env", environ=None):
    """DatabaseConnection::get_database_url: load `.env` (dotenv semantics: KEY=VALUE lines, `#` comments, optional quotes,
    variables already in the environment win) and return PGSQL_URL.  Raises where the reference panics."""
    import os
    environ = os.environ if environ is None else environ
    values = {}
    try:
        with open(env_path) as fh:
            for line in fh:
                line = line.strip()
                if not line or line.startswith("#") or "=" not in line:
                    continue
                if line.startswith("export "):
                    line = line[7:].lstrip()
                k, v = line.split("=", 1)
                v = v.strip()
                if len(v) >= 2 and v[0] == v[-1] and v[0] in "\"'":
                    v = v[1:-1]
                values[k.strip()] = v
    except OSError as err:
        raise RuntimeError("Could not load .env-file, reason: %s" % err)
    url = environ.get("PGSQL_URL", values.get("PGSQL_URL"))
    if not url:
        raise RuntimeError("Variable 'PGSQL_URL' must be set in .env")
    return url


TABLE_COLUMNS = {
    "proteins": "id, accession, header, aa_sequence, is_completely_digested",
    "peptides": "id, aa_sequence, length, number_of_missed_cleavages, weight, " + ", ".join(c + "_count" for c in COUNT_COLUMNS),
    "peptides_proteins": "peptide_id, protein_id",
    "decoys": "id, aa_sequence, length, number_of_missed_cleavages, weight, " + ", ".join(c + "_count" for c in COUNT_COLUMNS),
    "psms": "spectrum_id, scan_id, rank, is_decoy, peptide_id, aa_sequence, modres, weight, precursor_mass, charge, score, n_candidates",
}


def load_script(directory, tables=("proteins", "peptides", "peptides_proteins", "decoys", "psms"), env_path=None):
    """The psql script that loads the exported CSV files into the reference's schema (db/schema.sql) in dependency order --
    `\\copy` runs client side, so the files need not be readable by the server.  With `env_path` the command line that
    feeds it to the database of the reference's `.env` is returned too: (script, command)."""
    import os
    lines = ["-- load the tables exported by the B200 hot path into the max-decoy schema (db/schema.sql)", "BEGIN;"]
    if "psms" in tables:
        lines.append(PSMS_DDL.replace("CREATE TABLE psms", "CREATE TABLE IF NOT EXISTS psms").rstrip())
    for t in tables:
        path = os.path.join(directory, t + ".csv")
        lines.append("\\copy %s (%s) FROM '%s' WITH (FORMAT csv)" % (t, TABLE_COLUMNS[t], path))
    for t in ("proteins", "peptides", "decoys"):
        if t in tables:      # the serial ids were given explicitly: move the sequences behind them
            lines.append("SELECT setval(pg_get_serial_sequence('%s', 'id'), (SELECT COALESCE(MAX(id), 1) FROM %s));" % (t, t))
    lines.append("COMMIT;")
    script = "\n".join(lines) + "\n"
    if env_path is None:
        return script
    return script, ["psql", database_url(env_path), "-v", "ON_ERROR_STOP=1", "-f", os.path.join(directory, "load.sql")]


def unload_script(directory, tables=("peptides", "decoys")):
    """The other direction: `\\copy ... TO` for the tables the hot path takes as input (`peptides` for a digest done by the
    reference, `decoys` for md_decoy_store_set / --stored-decoys), in the column order the readers here expect."""
    import os
    return "".join("\\copy (SELECT %s FROM %s ORDER BY id) TO '%s' WITH (FORMAT csv)\n" % (TABLE_COLUMNS[t], t, os.path.join(directory, t + ".csv")) for t in tables)


def read_sequences_csv(text):
    """aa_sequence column (second field) of a `peptides` / `decoys` CSV as exported above or by unload_script."""
    return [row[1] for row in csv.reader(io.StringIO(text)) if len(row) > 1 and row[1] and row[1].isalpha()]
