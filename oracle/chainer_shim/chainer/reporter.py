last_report = {}


def report(values, observer=None):
    last_report.clear()
    last_report.update(values)
