"""Import-compatible stand-in for the reference's matplotlib wrapper (myutils/plot_helper.py).

Plotting is out of scope for the B200 hot path; the class only exists so that
`from myutils.plot_helper import PlotHelper` keeps working.  It needs matplotlib at construction time,
exactly like the reference (lazy import, plot_helper.py:23-30)."""


class PlotHelper:

    def __init__(self, axis_labels, fancy=False, font_size=12, ax=None, pgf=False):
        import matplotlib.pyplot as plt  # raises ImportError without matplotlib, as the reference does
        self.FONT_SIZE = font_size
        self.fig = None
        if ax is not None:
            self.ax = ax
        elif len(axis_labels) == 3:
            self.fig = plt.figure()
            self.ax = self.fig.add_subplot(projection='3d')
        else:
            self.fig, self.ax = plt.subplots()
        for setter, label in zip(('set_xlabel', 'set_ylabel', 'set_zlabel'), axis_labels):
            getattr(self.ax, setter)(label)

    def finalize(self, width=6, height=4, legendLoc=0, show_legend=True, **kw):
        if self.fig is not None:
            self.fig.set_size_inches(width, height)
        return self.ax.legend(loc=legendLoc) if show_legend else None

    def save(self, path):
        self.fig.savefig(path)

    def show(self):
        import matplotlib.pyplot as plt
        plt.show()
