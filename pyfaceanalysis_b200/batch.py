"""Batch mode of the reference (``python FaceDetect.py --batch=batch_filename``) around the device cascade.

What the reference does per image (``FaceDetectUpdated.py:513-1280``): open the file in mode 'L'
(``load_images([...], image_format="L")``, ``:533``), NEAREST prescale to at most 1000 pixels a side (``:551-559``),
the window pyramid and the cascade, eyes, purge, optionally age / race / gender, then APPEND one text line per face to
the image's output file (``:1258-1278``).  Here the files of a group are decoded by host threads (Pillow releases the GIL
while it decodes) while the previous group is on the GPU, a group is prescaled by one launch per image size and runs
through ``FaceDetector.detect`` as one batch, and the lines are written with ``cascade.format_detections``.  Coordinates
are those of the prescaled image, as in the reference (``prescaling_factor`` is not applied to the results there).

The decode itself stays on the host (Pillow, exactly the reference's pixels): it is outside the hot path of SURVEY.md
section 8 and byte-exact decoding of arbitrary JPEG / PNG files is the library's job.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .cascade import format_detections


def read_batch_file(batch_filename):
    """``face_analysis.py:224-243``: the file holds pairs of lines -- an input image filename, then the output filename
    for its detections; a trailing unpaired line is ignored (``len(lines) / 2`` pairs)."""
    with open(batch_filename, "r") as f:
        lines = f.readlines()
    image_filenames, output_filenames = [], []
    for i in range(len(lines) // 2):
        image_filenames.append(lines[2 * i].rstrip())
        output_filenames.append(lines[2 * i + 1].rstrip())
    return image_filenames, output_filenames


def load_images(image_filenames, image_format="L"):
    """The reference's ``load_images(filenames, image_format="L")`` + ``images[i].load()`` (``FaceDetectUpdated.py:533-534``)
    as contiguous uint8 arrays: (H, W) for 'L', (H, W, 3) for 'RGB'."""
    from PIL import Image
    out = []
    for name in image_filenames:
        with Image.open(name) as im:
            out.append(np.ascontiguousarray(im.convert(image_format)))
    return out


def run_batch(detector, batch, smallest_face=0.20, image_prescaling=True, prescale_size=1000, group=64, decode_threads=8,
              right_screen_eye_first=False, estimate_attributes=False, write_age_race_gender_confidence=True,
              write_results=True, benchmark=None):
    """Detect faces in every image of a batch file and append the result lines to the paired output files.

    detector: ``cascade.FaceDetector``; batch: the batch file name or ``(image_filenames, output_filenames)``.
    Defaults are the reference's (``FaceDetectUpdated.py:84-122``: smallest_face 0.20, prescaling on, 1000 pixels,
    results written, left screen eye first).  ``group`` images are processed as one device batch; the next group is
    decoded meanwhile.  ``write_age_race_gender_confidence`` only applies with ``estimate_attributes`` (the reference
    writes the four extra fields when the estimates exist).  Returns the per-image detection arrays (after the purge),
    in batch-file order."""
    image_filenames, output_filenames = read_batch_file(batch) if isinstance(batch, str) else batch
    if len(image_filenames) != len(output_filenames):
        raise ValueError("%d image filenames but %d output filenames" % (len(image_filenames), len(output_filenames)))
    if group < 1:
        raise ValueError("group must be at least 1")
    n = len(image_filenames)
    results = [None] * n
    if n == 0:
        return results
    starts = list(range(0, n, group))
    with ThreadPoolExecutor(max(1, decode_threads)) as pool:

        def decode(start):
            names = image_filenames[start:start + group]
            return list(pool.map(lambda nm: load_images([nm], "L")[0], names))
        with ThreadPoolExecutor(1) as ahead:                     # one group in flight behind the device
            pending = ahead.submit(decode, starts[0])
            for gi, start in enumerate(starts):
                images = pending.result()
                if gi + 1 < len(starts):
                    pending = ahead.submit(decode, starts[gi + 1])
                if image_prescaling:
                    images = detector.prescale(images, prescale_size)
                if estimate_attributes:
                    dets, attrs = detector.detect(images, smallest_face=smallest_face, benchmark=benchmark,
                                                  estimate_attributes=True)
                else:
                    dets, attrs = detector.detect(images, smallest_face=smallest_face, benchmark=benchmark), None
                for k, det in enumerate(dets):
                    results[start + k] = det
                    if not write_results:
                        continue
                    extra = None
                    if attrs is not None and write_age_race_gender_confidence:
                        extra = (attrs[k]["age"], attrs[k]["race"], attrs[k]["gender"])
                    with open(output_filenames[start + k], "a") as fd:      # 'a': FaceDetectUpdated.py:1260
                        fd.write(format_detections(det, right_screen_eye_first, extra))
    return results
