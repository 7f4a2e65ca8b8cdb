"""Metrics sink replacing the reference's visdom wrapper (utils/visualize.py:44-55 `vis.plot(name, y)`): the same
series names (`return`, `avg_return`, `success_rate`, main.py:130-160) appended to a JSONL stream, and exported as
per-series CSV files with the `xData,yData` columns of the reference's `get_vis_data` (main.py:587-623)."""
import json
import os
import time


class MetricsSink:
    def __init__(self, directory=None, name="run"):
        self.series = {}
        self.dir, self.name = directory, name
        self._fh = None
        if directory:
            os.makedirs(directory, exist_ok=True)
            self._fh = open(os.path.join(directory, name + ".jsonl"), "a")

    def plot(self, name, y, x=None, **extra):
        """vis.plot(name, y): one more point of series `name` (x defaults to the point's index, like visdom's append)"""
        pts = self.series.setdefault(name, [])
        x = len(pts) if x is None else x
        pts.append((x, float(y)))
        if self._fh:
            rec = {"t": time.time(), "name": name, "x": x, "y": float(y)}
            rec.update(extra)
            self._fh.write(json.dumps(rec) + "\n")
            self._fh.flush()

    def last(self, name, default=None):
        pts = self.series.get(name)
        return pts[-1][1] if pts else default

    def export_csv(self, directory=None):
        """main.py:615-621: one CSV per series, header xData,yData"""
        directory = directory or self.dir
        os.makedirs(directory, exist_ok=True)
        paths = []
        for name, pts in self.series.items():
            path = os.path.join(directory, "%s_%s.csv" % (self.name, name))
            with open(path, "w") as f:
                f.write("xData,yData\n")
                for x, y in pts:
                    f.write("%s,%s\n" % (x, y))
            paths.append(path)
        return paths

    def close(self):
        if self._fh:
            self._fh.close()
            self._fh = None
