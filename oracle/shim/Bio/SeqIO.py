"""Bio.SeqIO.parse(handle, 'fasta') stand-in: yields records with .id, .description, .seq.

Used by the reference at midas/run/snps.py:59, midas/run/genes.py:68 (rec.id = first
whitespace-delimited token of the header line; str(rec.seq); len(rec.seq))."""


class Seq(str):
    def upper(self):
        return Seq(str.upper(self))


class SeqRecord:
    def __init__(self, id, description, seq):
        self.id = id
        self.name = id
        self.description = description
        self.seq = Seq(seq)


def parse(handle, fmt):
    if fmt != "fasta":
        raise ValueError("shim supports only fasta")
    close = False
    if isinstance(handle, str):
        handle = open(handle)
        close = True
    title, chunks = None, []
    for line in handle:
        if isinstance(line, bytes):
            line = line.decode()
        if line.startswith(">"):
            if title is not None:
                yield SeqRecord(title.split(None, 1)[0] if title.split() else "", title, "".join(chunks))
            title, chunks = line[1:].rstrip("\r\n"), []
        elif title is not None:
            chunks.append(line.strip().replace(" ", ""))
    if title is not None:
        yield SeqRecord(title.split(None, 1)[0] if title.split() else "", title, "".join(chunks))
    if close:
        handle.close()
