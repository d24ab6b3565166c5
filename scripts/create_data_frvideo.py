#!/usr/bin/env python3
"""Turn a video file into a DAVIS2016-style dataset folder so `test_generator.py --dataset=DAVIS2016 --root_dir=<out>` can run on it
(role of the reference's scripts/create_data_frvideo.py, which shells out to ffmpeg; this one decodes with OpenCV).

    <out>/JPEGImages/480p/<name>/00000.jpg ...     frames resampled to --fps, resized to 853x480
    <out>/Annotations/480p/<name>/00000.png        one all-black mask every frame line points to (no ground truth for a raw video)
    <out>/ImageSets/480p/val.txt                   '/JPEGImages/480p/<name>/NNNNN.jpg /Annotations/480p/<name>/00000.png' per frame

Usage: python scripts/create_data_frvideo.py VIDEO [--out DIR] [--fps 24] [--size 853x480] [--max_frames N]"""
import argparse
import os

import cv2
import numpy as np


def convert(video, out, fps=24.0, size=(853, 480), max_frames=None):
    cap = cv2.VideoCapture(video)
    if not cap.isOpened():
        raise IOError("Could not open video %s" % video)
    name = os.path.splitext(os.path.basename(video))[0]
    img_dir = os.path.join(out, 'JPEGImages', '480p', name)
    ann_dir = os.path.join(out, 'Annotations', '480p', name)
    set_dir = os.path.join(out, 'ImageSets', '480p')
    for d in (img_dir, ann_dir, set_dir):
        os.makedirs(d, exist_ok=True)
    src_fps = cap.get(cv2.CAP_PROP_FPS) or fps
    step = max(src_fps / float(fps), 1e-6)       # source frames per output frame (>= keeps every frame when the source is slower)
    lines, k, next_pick, idx = [], 0, 0.0, 0
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        if idx + 1e-9 >= next_pick:
            frame = cv2.resize(frame, size, interpolation=cv2.INTER_AREA)
            cv2.imwrite(os.path.join(img_dir, '%05d.jpg' % k), frame, [cv2.IMWRITE_JPEG_QUALITY, 95])
            lines.append('/JPEGImages/480p/%s/%05d.jpg /Annotations/480p/%s/00000.png' % (name, k, name))
            k += 1
            next_pick += step
            if max_frames and k >= max_frames:
                break
        idx += 1
    cap.release()
    if k == 0:
        raise IOError("No frames decoded from %s" % video)
    cv2.imwrite(os.path.join(ann_dir, '00000.png'), np.zeros((size[1], size[0]), np.uint8))
    with open(os.path.join(set_dir, 'val.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    return k


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('video')
    ap.add_argument('--out', default=None, help='dataset root to create (default: <video folder>)')
    ap.add_argument('--fps', type=float, default=24.0, help='output frame rate (DAVIS 2016 is 24 fps)')
    ap.add_argument('--size', default='853x480')
    ap.add_argument('--max_frames', type=int, default=None)
    a = ap.parse_args()
    w, h = (int(v) for v in a.size.split('x'))
    out = a.out or os.path.dirname(os.path.abspath(a.video))
    n = convert(a.video, out, a.fps, (w, h), a.max_frames)
    print('wrote %d frames under %s' % (n, out))


if __name__ == '__main__':
    main()
