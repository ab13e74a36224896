"""Input preparation and result output around gat_b200.run() (reference: gat/IO.py, BED reader in
gat/Engine.pyx:2470-2556).  One-off host work, O(input); not part of the simulated hot path.
"""
import collections
import glob
import gzip
import os
import re
import sys

import numpy as np

from . import engine as Engine
from .segmentlist import SegmentList


def openFile(filename, mode="r"):
    if filename.endswith(".gz"):
        return gzip.open(filename, mode + "t")
    return open(filename, mode)


def _track_name(line):
    """name=... attribute of a BED `track` line"""
    m = re.search(r'name=("([^"]*)"|(\S+))', line)
    if not m:
        return None
    return m.group(2) if m.group(2) is not None else m.group(3)


def _readBedColumns(filename):
    """fast path for plain BED files (no `track` / comment lines, a constant number of >= 3 tab-separated
    columns): -> ((contig codes, contig names), start int64, end int64, (name codes, names) or None), or
    None when the file needs the line-by-line reader.  Multi-threaded Arrow CSV reader with dictionary-encoded
    string columns; the per-line Python loop costs ~1 us/line, i.e. ~25 s for the 2e7 intervals of the
    1000-track shape."""
    try:
        import pyarrow as pa
        import pyarrow.csv as pacsv
        with openFile(filename, "r") as f:
            first = f.readline()
        ncol = len(first.rstrip("\n").split("\t"))
        if ncol < 3 or first.startswith(("track", "#")) or not first.strip():
            return None
        names = ["c%i" % i for i in range(ncol)]
        dic = pa.dictionary(pa.int32(), pa.string())
        types = {"c0": dic, "c1": pa.int64(), "c2": pa.int64()}
        if ncol > 3:
            types["c3"] = dic
        table = pacsv.read_csv(
            filename, read_options=pacsv.ReadOptions(column_names=names, block_size=1 << 24),
            parse_options=pacsv.ParseOptions(delimiter="\t", quote_char=False),
            convert_options=pacsv.ConvertOptions(column_types=types, include_columns=list(types),
                                                 strings_can_be_null=False))

        def codes(col):
            arr = table[col].unify_dictionaries().combine_chunks()
            # (a zero-copy view: `zero_copy_only=False` would import pandas, 1.5-3 s)
            return np.frombuffer(arr.indices.buffers()[1], dtype=np.int32, count=len(arr), offset=arr.indices.offset * 4
                                 ).astype(np.int64), arr.dictionary.to_pylist()

        contig = codes("c0")
        if any(c.startswith(("track", "#")) for c in contig[1]):
            return None
        name = codes("c3") if ncol > 3 else None
        def column(col):
            a = table[col].combine_chunks()
            if a.null_count:
                raise ValueError("empty coordinate")
            return np.frombuffer(a.buffers()[1], dtype=np.int64, count=len(a), offset=a.offset * 8)

        return contig, column("c1"), column("c2"), name
    except Exception:
        return None


def _firstAppearance(codes, n):
    """relabel codes 0..n-1 in order of first appearance -> (new codes, old label of each new code)"""
    N = len(codes)
    first = np.full(n, N, dtype=np.int64)
    # assigning the row numbers in REVERSE order leaves every code with its smallest row (the last write wins);
    # np.minimum.at does the same 20x slower
    first[codes[::-1]] = np.arange(N - 1, -1, -1, dtype=np.int64)
    order = np.argsort(first, kind="stable")
    order = order[first[order] < N]
    remap = np.zeros(n, dtype=np.int64)
    remap[order] = np.arange(len(order))
    return remap[codes], order


def _groupRows(result_rows, origin, filename, default_name, cols, allow_multiple, ignore_tracks):
    """split the columns of one file into rows[track][contig] (first-appearance order of tracks and contigs)"""
    (ccodes, cnames), start, end, name = cols
    n = len(ccodes)
    if ignore_tracks or name is None:
        tcodes, tnames = np.zeros(n, dtype=np.int64), ["merged" if ignore_tracks else default_name]
    else:
        tcodes, order = _firstAppearance(name[0], len(name[1]))
        tnames = [name[1][i] if name[1][i] else default_name for i in order]
    for t in tnames:
        if t in origin and origin[t] != filename and not allow_multiple:
            raise ValueError("track '%s' in multiple filenames: %s and %s" % (t, origin[t], filename))
        origin[t] = filename
    key = tcodes * len(cnames) + ccodes
    order = np.argsort(key, kind="stable")
    key = key[order]
    pairs = np.stack([start[order], end[order]], axis=1)
    bounds = np.flatnonzero(np.diff(key)) + 1
    firsts = np.concatenate([[0], bounds]).astype(np.int64)
    lasts = np.concatenate([bounds, [len(key)]]).astype(np.int64)
    # contigs of a track in order of first appearance in the file, like the line-by-line reader
    first_line = np.minimum.reduceat(order, firsts) if len(key) else np.zeros(0, dtype=np.int64)
    for g in np.lexsort((first_line, key[firsts] // len(cnames))):
        a, b = firsts[g], lasts[g]
        result_rows[tnames[key[a] // len(cnames)]][cnames[key[a] % len(cnames)]].append(pairs[a:b])


def _readFlat(filenames, allow_multiple, ignore_tracks):
    """all files through the column reader -> one table of rows (track index, contig index, start, end) with tracks
    and contigs numbered in order of first appearance; None when a file needs the line-by-line reader or holds a
    coordinate the device path does not take (negative, >= 2^31)"""
    tracks, contigs, origin, per_file = {}, {}, {}, []
    parts = []
    for filename in filenames:
        cols = _readBedColumns(filename)
        if cols is None:
            return None
        (ccodes, cnames), start, end, name = cols
        if len(start) and (int(start.min()) < 0 or int(end.min()) < 0 or int(start.max()) >= 2 ** 31 or int(end.max()) >= 2 ** 31
                           or bool((start > end).any())):
            return None
        default_name = os.path.basename(filename)
        n = len(ccodes)
        if ignore_tracks or name is None:
            tcodes, tnames = np.zeros(n, dtype=np.int64), ["merged" if ignore_tracks else default_name]
        else:
            tcodes, order = _firstAppearance(name[0], len(name[1]))
            tnames = [name[1][i] if name[1][i] else default_name for i in order]
        for t in tnames:
            if t in origin and origin[t] != filename and not allow_multiple:
                raise ValueError("track '%s' in multiple filenames: %s and %s" % (t, origin[t], filename))
            origin[t] = filename
        ccodes2, corder = _firstAppearance(ccodes, len(cnames))
        tmap = np.array([tracks.setdefault(t, len(tracks)) for t in tnames], dtype=np.int64)
        cmap = np.array([contigs.setdefault(cnames[i], len(contigs)) for i in corder], dtype=np.int64)
        parts.append((tmap[tcodes] if n else tcodes, cmap[ccodes2] if n else ccodes2, start, end))
        per_file.append((filename, default_name, cols))

    def grouper():          # the host reader's grouping of the same rows (only if the collection falls back to the host)
        rows = collections.defaultdict(lambda: collections.defaultdict(list))
        seen = {}
        for filename, default_name, cols in per_file:
            _groupRows(rows, seen, filename, default_name, cols, True, ignore_tracks)
        return rows

    cat = (lambda i: np.concatenate([p[i] for p in parts]) if len(parts) > 1 else parts[0][i]) if parts else None
    if not parts:
        return None
    return dict(tracks=list(tracks), contigs=list(contigs), track=cat(0), contig=cat(1), start=cat(2), end=cat(3),
                grouper=grouper)


def readFromBed(filenames, allow_multiple=False, ignore_tracks=False, flat=False):
    """BED files -> {track: IntervalDictionary}.  The track of an interval is the enclosing `track
    name=` line, else column 4, else the file's basename; `ignore_tracks` pools everything into
    'merged' (gat/Engine.pyx:2470-2556).  flat: -> a table of rows for Engine.DeviceIntervalCollection (see
    _readFlat) when every file can be read by columns, else the dictionary as usual."""
    if isinstance(filenames, str):
        filenames = [filenames]
    if flat:
        table = _readFlat(filenames, allow_multiple, ignore_tracks)
        if table is not None:
            return table
    rows = collections.defaultdict(lambda: collections.defaultdict(list))
    origin = {}
    for filename in filenames:
        default_name = os.path.basename(filename)
        cols = _readBedColumns(filename)
        if cols is not None:
            _groupRows(rows, origin, filename, default_name, cols, allow_multiple, ignore_tracks)
            continue
        current = None
        chunk = collections.defaultdict(lambda: collections.defaultdict(list))
        with openFile(filename, "r") as infile:
            for lineno, line in enumerate(infile):
                if line.startswith("track"):
                    current = _track_name(line)
                    if current is None:
                        raise KeyError("track without field 'name' in file '%s'" % filename)
                    continue
                if line.startswith("#") or not line.strip():
                    continue
                fields = line.rstrip("\n").split("\t")
                if len(fields) < 3:
                    raise IOError("malformatted entry in line %s:%i" % (filename, lineno))
                if ignore_tracks:
                    name = "merged"
                elif current is not None:
                    name = current
                elif len(fields) > 3 and fields[3]:
                    name = fields[3]
                else:
                    name = default_name
                if name in origin:
                    if origin[name] != filename:
                        if not allow_multiple:
                            raise ValueError("track '%s' in multiple filenames: %s and %s" %
                                             (name, origin[name], filename))
                        origin[name] = filename
                else:
                    origin[name] = filename
                chunk[name][fields[0]].append((int(fields[1]), int(fields[2])))
        for name, contigs in chunk.items():
            for contig, data in contigs.items():
                rows[name][contig].append(np.array(data, dtype=np.int64).reshape(-1, 2))
    result = collections.defaultdict(Engine.IntervalDictionary)
    for name, contigs in rows.items():
        d = Engine.IntervalDictionary()
        for contig, parts in contigs.items():
            data = parts[0] if len(parts) == 1 else np.concatenate(parts)
            d[contig] = SegmentList(array=data.astype(np.int64).astype(np.uint32))
        result[name] = d
    return result


def readSegmentList(label, filenames, enable_split_tracks=False, ignore_tracks=False, on_device=False):
    """on_device: keep the lists on the GPU (Engine.DeviceIntervalCollection) -- for the large annotation collection"""
    results = Engine.DeviceIntervalCollection(name=label) if on_device else Engine.IntervalCollection(name=label)
    results.load(filenames, allow_multiple=enable_split_tracks, ignore_tracks=ignore_tracks)
    return results


def expandGlobs(infiles):
    out = []
    for x in infiles:
        out.extend(glob.glob(x))
    return out


def buildSegments(options):
    """load segments, annotations, workspace and isochores as named by *options*
    (gat/IO.py:88-185): normalize, collapse the workspaces into one, truncate isochores to it."""
    options.segment_files = expandGlobs(options.segment_files)
    options.annotation_files = expandGlobs(options.annotation_files)
    options.workspace_files = expandGlobs(options.workspace_files)
    if not options.segment_files:
        raise ValueError("please specify at least one segment file")
    if not options.annotation_files:
        raise ValueError("please specify at least one annotation file")
    if not options.workspace_files:
        raise ValueError("please specify at least one workspace file")

    segments = readSegmentList("segments", options.segment_files, ignore_tracks=options.ignore_segment_tracks)
    segments.normalize()
    if segments.sum() == 0:
        raise ValueError("segments file is empty - run aborted")
    if len(segments) > 1000:
        raise ValueError("too many (%i) segment files - use track definitions or --ignore-segment-tracks"
                         % len(segments))

    # the annotations are the large input (1000 tracks x 20 000 intervals): their preparation -- normalize, the
    # intersection with the workspace, the isochore split -- runs on the GPU (GATB_HOST_PREP=1: on the host)
    annotations = readSegmentList("annotations", options.annotation_files,
                                  enable_split_tracks=options.enable_split_tracks,
                                  ignore_tracks=options.annotations_label is not None,
                                  on_device=not os.environ.get("GATB_HOST_PREP"))
    if options.annotations_label is not None:
        annotations.setName(options.annotations_label)
    if getattr(options, "annotations_to_points", None):
        raise NotImplementedError("--annotations-to-points is not accelerated by gat_b200")
    if getattr(options, "overlapping_annotations", False):
        raise NotImplementedError("--overlapping-annotations is not accelerated by gat_b200")
    annotations.normalize()

    workspaces = readSegmentList("workspaces", options.workspace_files, enable_split_tracks=True,
                                 ignore_tracks=options.enable_split_tracks)
    workspaces.normalize()
    workspaces.collapse()
    workspaces.restrict("collapsed")

    isochores = None
    if options.isochore_files:
        isochores = Engine.IntervalCollection(name="isochores")
        isochores.load(options.isochore_files)
        isochores.sort()
        for s in isochores.getSegmentLists():
            s.check()
        isochores.normalize()
        isochores.intersect(workspaces["collapsed"])
    return segments, annotations, workspaces, isochores


def applyIsochores(segments, annotations, workspaces, options, isochores=None,
                   truncate_segments_to_workspace=False, truncate_workspace_to_annotations=False,
                   restrict_workspace=False):
    """restrict segments/annotations to the workspace and optionally split everything by isochore
    (gat/IO.py:188-293); returns the workspace IntervalDictionary."""
    if isochores:
        workspaces.toIsochores(isochores, truncate=True)
        annotations.toIsochores(isochores, truncate=True)
        segments.toIsochores(isochores, truncate=options.truncate_segments_to_workspace)
        if workspaces.sum() == 0:
            raise ValueError("isochores and workspaces do not overlap")
        if annotations.sum() == 0:
            raise ValueError("isochores and annotations do not overlap")
        if segments.sum() == 0:
            raise ValueError("isochores and segments do not overlap")
    else:
        if options.truncate_segments_to_workspace:
            segments.intersect(workspaces["collapsed"])
        else:
            segments.filter(workspaces["collapsed"])
        annotations.intersect(workspaces["collapsed"])

    workspace = workspaces["collapsed"]
    if restrict_workspace:
        for _ in (segments, annotations):
            if "merged" in segments:
                workspace.filter(segments["merged"])
            else:
                segments.merge()
                workspace.filter(segments["merged"])
                del segments["merged"]
    if truncate_workspace_to_annotations:
        annotations.merge()
        annotations["merged"].normalize()
        workspace.intersect(annotations["merged"])
        del annotations["merged"]
    return workspace


def readDescriptions(options):
    """optional table annotation -> description columns (gat/IO.py:296-330)"""
    header, descriptions, width = [], {}, 0
    fn = getattr(options, "input_filename_descriptions", None)
    if fn:
        with openFile(fn) as inf:
            first = True
            for line in inf:
                if line.startswith("#"):
                    continue
                data = line.rstrip("\n").split("\t")
                if first:
                    header = data[1:]
                    width = len(header)
                    first = False
                    continue
                descriptions[data[0]] = data[1:]
    return header, descriptions, width


def outputResults(results, options, header, description_header, description_width, descriptions,
                  format_observed="%i"):
    """q-values over ALL results of the run, then one table per counter (gat/IO.py:457-538)"""
    pvalues = [x.pvalue for x in results]
    qvalues = Engine.getQValues(pvalues, method=options.qvalue_method, vlambda=options.qvalue_lambda,
                                pi0_method=options.qvalue_pi0_method)
    for x, q in zip(results, qvalues):
        x.qvalue = q
        x.format_observed = format_observed

    counters = sorted(set(x.counter for x in results))
    keyfuncs = {"track": lambda x: (x.track, x.annotation), "observed": lambda x: x.observed,
                "annotation": lambda x: (x.annotation, x.track), "fold": lambda x: x.fold,
                "pvalue": lambda x: x.pvalue, "qvalue": lambda x: x.qvalue}
    if options.output_order not in keyfuncs:
        raise ValueError("unknown sort order %s" % options.output_order)
    stdout = getattr(options, "stdout", sys.stdout)
    for counter in counters:
        if len(counters) == 1:
            outfile, output = stdout, list(results)
        else:
            outfile = openFile(re.sub("%s", counter, options.output_tables_pattern), "w")
            output = [x for x in results if x.counter == counter]
        outfile.write("\t".join(list(header) + list(description_header)) + "\n")
        output.sort(key=keyfuncs[options.output_order])
        for result in output:
            outfile.write(str(result))
            if descriptions:
                outfile.write("\t" + "\t".join(descriptions.get(result.annotation, [""] * description_width)))
            outfile.write("\n")
        if outfile is not stdout:
            outfile.close()
