"""Drawing for the README API (``results.show()`` / ``results.save()``, README.md:48-53).

Interface of the reference's ``ImageAnnotator`` (skyeye/utils/visualization.py:43-147: ``box_label(box, label, color,
txt_color)``, ``result()``), OpenCV backend (the reference module itself cannot be imported: it needs matplotlib,
seaborn and an undefined ``Path``).  Host-side, after the hot path."""
from __future__ import annotations

import numpy as np

_PALETTE = ("FF3838", "FF9D97", "FF701F", "FFB21D", "CFD231", "48F90A", "92CC17", "3DDB86", "1A9334", "00D4BB",
            "2C99A8", "00C2FF", "344593", "6473FF", "0018EC", "8438FF", "520085", "CB38FF", "FF95C8", "FF37C7")


def colors(i: int, bgr: bool = True):
    h = _PALETTE[int(i) % len(_PALETTE)]
    rgb = tuple(int(h[k:k + 2], 16) for k in (0, 2, 4))
    return rgb[::-1] if bgr else rgb


class ImageAnnotator:
    def __init__(self, image: np.ndarray, line_width=None, font_size=None, font="Arial.ttf", pil=False, example="abc"):
        if pil:
            raise NotImplementedError("the PIL backend of the reference annotator is not provided; OpenCV draws ASCII labels")
        self.im = np.ascontiguousarray(image)
        self.line_width = line_width or max(round(sum(self.im.shape) / 2 * 0.003), 2)

    def box_label(self, box, label="", color=(128, 128, 128), txt_color=(255, 255, 255)):
        """One xyxy box with an optional filled label tag above it (inside when it does not fit), visualization.py:101-118."""
        import cv2
        p1, p2 = (int(box[0]), int(box[1])), (int(box[2]), int(box[3]))
        cv2.rectangle(self.im, p1, p2, color, thickness=self.line_width, lineType=cv2.LINE_AA)
        if label:
            tf = max(self.line_width - 1, 1)
            fs = self.line_width / 3
            w, h = cv2.getTextSize(label, 0, fontScale=fs, thickness=tf)[0]
            outside = p1[1] - h - 3 >= 0
            q = (p1[0] + w, p1[1] - h - 3 if outside else p1[1] + h + 3)
            cv2.rectangle(self.im, p1, q, color, -1, cv2.LINE_AA)
            cv2.putText(self.im, label, (p1[0], p1[1] - 2 if outside else p1[1] + h + 2), 0, fs, txt_color, thickness=tf,
                        lineType=cv2.LINE_AA)

    def result(self) -> np.ndarray:
        return np.asarray(self.im)
